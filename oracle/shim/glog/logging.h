// TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.
// Minimal stand-in for the glog macros the reference's sources use (LOG, VLOG, CHECK*), so that they compile in
// an image without glog.  INFO/VLOG output is dropped unless HPMVS_REF_VERBOSE is set; WARNING/ERROR go to stderr;
// a failed CHECK aborts like glog's.
#ifndef HPMVS_ORACLE_GLOG_SHIM_H
#define HPMVS_ORACLE_GLOG_SHIM_H
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace google {
enum { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
inline void InitGoogleLogging(const char*) {}
inline bool shim_verbose() { static const bool v = std::getenv("HPMVS_REF_VERBOSE") != nullptr; return v; }
class LogMessage {
public:
    LogMessage(int sev, bool on) : sev_(sev), on_(on) {}
    ~LogMessage() {
        if (on_) { ss_ << "\n"; std::cerr << ss_.str(); }
        if (sev_ == FATAL) std::abort();
    }
    std::ostream& stream() { return ss_; }
private:
    int sev_; bool on_; std::ostringstream ss_;
};
struct Voidify { void operator&(std::ostream&) {} };
}  // namespace google

static bool FLAGS_logtostderr = false, FLAGS_colorlogtostderr = false;

#define HPMVS_SHIM_SEV_INFO google::INFO
#define HPMVS_SHIM_SEV_WARNING google::WARNING
#define HPMVS_SHIM_SEV_ERROR google::ERROR
#define HPMVS_SHIM_SEV_FATAL google::FATAL
#define LOG(sev) google::LogMessage(HPMVS_SHIM_SEV_##sev, HPMVS_SHIM_SEV_##sev != google::INFO || google::shim_verbose()).stream()
#define VLOG(n) google::LogMessage(google::INFO, false).stream()
#define CHECK(cond) (cond) ? (void)0 : google::Voidify() & google::LogMessage(google::FATAL, true).stream() << "Check failed: " #cond " "
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
template <class T> inline T* hpmvs_shim_check_notnull(T* p, const char* what) {
    if (!p) { std::cerr << "Check failed: '" << what << "' Must be non NULL\n"; std::abort(); }
    return p;
}
#define CHECK_NOTNULL(p) hpmvs_shim_check_notnull((p), #p)
#endif
