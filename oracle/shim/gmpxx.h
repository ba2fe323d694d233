// oracle/shim/gmpxx.h -- TEST INFRASTRUCTURE ONLY.
// The image has the GMP runtime (libgmp.so.10, 6.3.0) but no headers.  This is the small
// slice of mpz_class the reference TU uses (NDP-5_6_7.cpp:997-1027,1074-1077,1130-1153,
// 1609-1617), written over hand-declared __gmpz_* prototypes.  Link with -l:libgmp.so.10.
#ifndef NDP_ORACLE_SHIM_GMPXX_H
#define NDP_ORACLE_SHIM_GMPXX_H
#include <cstdlib>
#include <ostream>
#include <string>

extern "C" {
typedef struct { int _mp_alloc; int _mp_size; unsigned long* _mp_d; } ndp_shim_mpz_struct;
void __gmpz_init(ndp_shim_mpz_struct*);
void __gmpz_clear(ndp_shim_mpz_struct*);
void __gmpz_set(ndp_shim_mpz_struct*, const ndp_shim_mpz_struct*);
void __gmpz_set_si(ndp_shim_mpz_struct*, long);
int __gmpz_set_str(ndp_shim_mpz_struct*, const char*, int);
char* __gmpz_get_str(char*, int, const ndp_shim_mpz_struct*);
void __gmpz_mul(ndp_shim_mpz_struct*, const ndp_shim_mpz_struct*, const ndp_shim_mpz_struct*);
void __gmpz_mul_si(ndp_shim_mpz_struct*, const ndp_shim_mpz_struct*, long);
void __gmpz_add_ui(ndp_shim_mpz_struct*, const ndp_shim_mpz_struct*, unsigned long);
void __gmpz_sub_ui(ndp_shim_mpz_struct*, const ndp_shim_mpz_struct*, unsigned long);
int __gmpz_cmp(const ndp_shim_mpz_struct*, const ndp_shim_mpz_struct*);
}

class mpz_class {
  public:
    mpz_class() { __gmpz_init(&z_); }
    mpz_class(int v) { __gmpz_init(&z_); __gmpz_set_si(&z_, v); }
    mpz_class(const mpz_class& o) { __gmpz_init(&z_); __gmpz_set(&z_, &o.z_); }
    ~mpz_class() { __gmpz_clear(&z_); }
    mpz_class& operator=(const mpz_class& o) { __gmpz_set(&z_, &o.z_); return *this; }
    int set_str(const std::string& s, int base) { return __gmpz_set_str(&z_, s.c_str(), base); }
    std::string get_str(int base = 10) const {
        char* p = __gmpz_get_str(nullptr, base, &z_);
        std::string s(p);
        std::free(p);
        return s;
    }
    mpz_class& operator*=(int v) { __gmpz_mul_si(&z_, &z_, v); return *this; }
    mpz_class& operator+=(int v) {
        if (v >= 0) __gmpz_add_ui(&z_, &z_, (unsigned long)v);
        else __gmpz_sub_ui(&z_, &z_, (unsigned long)(-(long)v));
        return *this;
    }
    friend mpz_class operator*(const mpz_class& a, const mpz_class& b) {
        mpz_class r; __gmpz_mul(&r.z_, &a.z_, &b.z_); return r;
    }
    friend bool operator==(const mpz_class& a, const mpz_class& b) { return __gmpz_cmp(&a.z_, &b.z_) == 0; }
    friend std::ostream& operator<<(std::ostream& os, const mpz_class& a) { return os << a.get_str(); }
  private:
    ndp_shim_mpz_struct z_;
};
#endif
