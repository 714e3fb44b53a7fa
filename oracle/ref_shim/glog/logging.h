// ORACLE — TEST INFRASTRUCTURE ONLY.  glog stand-in: CHECKs abort with a message, LOG() swallows its stream.
#pragma once
#include <cstdio>
#include <cstdlib>
namespace vshim {
struct NullStream {
  template <class T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
inline void check_failed(const char* what, const char* file, int line) {
  std::fprintf(stderr, "Check failed: %s (%s:%d)\n", what, file, line);
  std::abort();
}
template <class T> T& check_notnull(T& p, const char* what, const char* file, int line) {
  if (p == nullptr) check_failed(what, file, line);
  return p;
}
}  // namespace vshim
#include <ostream>
#define LOG(sev) ::vshim::NullStream()
#define VLOG(n) ::vshim::NullStream()
#define LOG_IF(sev, c) ::vshim::NullStream()
#define CHECK(c) ((c) ? (void)0 : ::vshim::check_failed(#c, __FILE__, __LINE__)), ::vshim::NullStream()
#define CHECK_OP_(a, op, b) (((a)op(b)) ? (void)0 : ::vshim::check_failed(#a " " #op " " #b, __FILE__, __LINE__)), ::vshim::NullStream()
#define CHECK_GT(a, b) CHECK_OP_(a, >, b)
#define CHECK_GE(a, b) CHECK_OP_(a, >=, b)
#define CHECK_LT(a, b) CHECK_OP_(a, <, b)
#define CHECK_LE(a, b) CHECK_OP_(a, <=, b)
#define CHECK_EQ(a, b) CHECK_OP_(a, ==, b)
#define CHECK_NE(a, b) CHECK_OP_(a, !=, b)
#define CHECK_NOTNULL(p) ::vshim::check_notnull(p, #p, __FILE__, __LINE__)
