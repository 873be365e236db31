// TEST INFRASTRUCTURE ONLY -- run-time support of the C++ that oracle/f2cpp.py generates from
// the reference's Fortran sources (oracle/_ref/, never committed, never linked into the product).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

static std::string f2c_last_error;

// static storage of the Fortran program: zero-initialised, like gfortran's .bss
template <class T>
static T *f2c_alloc(long n) {
  if (n < 0) n = 0;
  void *p = std::calloc((size_t)n + 1, sizeof(T));
  if (!p) throw std::runtime_error("f2c_alloc: out of memory");
  return static_cast<T *>(p);
}

// an expression passed as an actual argument: Fortran passes the address of a temporary
template <class T>
static T *f2c_tmp_slot(const T &v) {
  static thread_local T ring[64];
  static thread_local int at = 0;
  T *p = &ring[at++ & 63];
  *p = v;
  return p;
}
template <class T, class U>
static T *f2c_tmp(const U &v) {
  return f2c_tmp_slot<T>((T)v);
}

// x**n with integer n by repeated multiplication (gfortran expands small constant powers the
// same way; for the reference's uses -- **2, **3, **4 -- the result is the product chain)
template <class T>
static T f2c_ipow(T x, int n) {
  if (n < 0) return (T)(1.0 / (double)f2c_ipow<T>(x, -n));
  T r = (T)1;
  while (n > 0) {
    if (n & 1) r = r * x;
    n >>= 1;
    if (n) x = x * x;
  }
  return r;
}

static inline double f2c_sign(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }
static inline int f2c_isign(int a, int b) { return b < 0 ? -std::abs(a) : std::abs(a); }
static inline std::string f2c_trim(const std::string &s) {
  size_t e = s.find_last_not_of(' ');
  return e == std::string::npos ? std::string() : s.substr(0, e + 1);
}
[[noreturn]] static inline void f2c_stop(const char *where) { throw std::runtime_error(std::string("STOP at ") + where); }
