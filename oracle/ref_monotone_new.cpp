// TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.
//
// Monotone global operator new for the two command-line builds that tests/test_dropin.py compares (hpmvs_ref_det,
// hpmvs_ref_b200_det).  Why: the reference's processing order depends on HEAP ADDRESSES - CellProcessor::branch collects
// the new cells in a std::set<Leaf<Ppatch3d>*> and queues them in pointer order (src/hpmvs/CellProcessor.cpp:289-305),
// and equal-priority cells keep their insertion order - so two different binaries (different temporaries, different
// malloc free lists) walk the cells in different orders and end with different patch sets although every optimize() call
// agrees bit for bit.  With addresses that only ever grow, pointer order = creation order in BOTH binaries, and their
// outputs become comparable byte for byte.  Nothing is ever freed: only for small test scenes.
#include <atomic>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <sys/mman.h>

namespace {
const size_t kReserve = size_t(64) << 30;   // address space only (MAP_NORESERVE)
char* arena() {
    static char* base = [] {
        void* p = mmap(nullptr, kReserve, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("ref_monotone_new: mmap"); abort(); }
        return static_cast<char*>(p);
    }();
    return base;
}
std::atomic<size_t> g_off{0};
void* bump(size_t n) {
    n = (n + 63) & ~size_t(63);
    const size_t o = g_off.fetch_add(n);
    if (o + n > kReserve) { fprintf(stderr, "ref_monotone_new: arena exhausted\n"); abort(); }
    return arena() + o;
}
}  // namespace

void* operator new(size_t n) { return bump(n); }
void* operator new[](size_t n) { return bump(n); }
void* operator new(size_t n, const std::nothrow_t&) noexcept { return bump(n); }
void* operator new[](size_t n, const std::nothrow_t&) noexcept { return bump(n); }
void operator delete(void*) noexcept {}
void operator delete[](void*) noexcept {}
void operator delete(void*, size_t) noexcept {}
void operator delete[](void*, size_t) noexcept {}
