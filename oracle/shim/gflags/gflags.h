// TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.
// Minimal stand-in for gflags (DEFINE_* + ParseCommandLineFlags with --name=value / --name value / --[no]name) so that
// the reference's src/main.cpp compiles in an image without gflags.
#ifndef HPMVS_ORACLE_GFLAGS_SHIM_H
#define HPMVS_ORACLE_GFLAGS_SHIM_H
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>

namespace gflags {
struct FlagRef { int kind; void* ptr; };   // 0 bool, 1 int32, 2 string
inline std::map<std::string, FlagRef>& registry() { static std::map<std::string, FlagRef> r; return r; }
struct Registrar { Registrar(const char* n, int k, void* p) { registry()[n] = FlagRef{k, p}; } };
inline void set_flag(const FlagRef& f, const std::string& v) {
    if (f.kind == 0) *static_cast<bool*>(f.ptr) = !(v == "false" || v == "0" || v == "no");
    else if (f.kind == 1) *static_cast<int*>(f.ptr) = std::atoi(v.c_str());
    else *static_cast<std::string*>(f.ptr) = v;
}
inline unsigned ParseCommandLineFlags(int* argc, char*** argv, bool) {
    for (int i = 1; i < *argc; i++) {
        std::string a = (*argv)[i];
        if (a.rfind("--", 0) == 0) a = a.substr(2); else if (a.rfind("-", 0) == 0) a = a.substr(1); else continue;
        std::string name = a, val; bool has = false;
        const size_t eq = a.find('=');
        if (eq != std::string::npos) { name = a.substr(0, eq); val = a.substr(eq + 1); has = true; }
        auto it = registry().find(name);
        if (it == registry().end() && name.rfind("no", 0) == 0) {
            it = registry().find(name.substr(2));
            if (it != registry().end() && it->second.kind == 0) { set_flag(it->second, "false"); continue; }
        }
        if (it == registry().end()) { std::cerr << "unknown flag --" << name << "\n"; std::exit(1); }
        if (!has) {
            if (it->second.kind == 0) val = "true";
            else if (i + 1 < *argc) val = (*argv)[++i];
        }
        set_flag(it->second, val);
    }
    return 1;
}
}  // namespace gflags
namespace google { using gflags::ParseCommandLineFlags; }
#ifndef HPMVS_GFLAGS_NAMESPACE
#define HPMVS_GFLAGS_NAMESPACE gflags
#endif
#define DEFINE_bool(name, def, help) bool FLAGS_##name = def; static gflags::Registrar reg_##name(#name, 0, &FLAGS_##name)
#define DEFINE_int32(name, def, help) int FLAGS_##name = def; static gflags::Registrar reg_##name(#name, 1, &FLAGS_##name)
#define DEFINE_string(name, def, help) std::string FLAGS_##name = def; static gflags::Registrar reg_##name(#name, 2, &FLAGS_##name)
#endif
