// Host build of the device arithmetic headers (portable primitive bodies) for GPU-less unit tests.
// TEST BUILD ONLY: the shipped library (libkgv.so) never contains these host paths.
#include "../../rusty_kaspa_b200/csrc/kgv_arith.cuh"
#include <cstring>
using namespace kgv;
static void load(fe& r, const uint8_t* b) { for (int i = 0; i < 8; i++) { uint32_t w; memcpy(&w, b + 4 * i, 4); r.v[i] = w; } }
static void store(uint8_t* b, const fe& r) { for (int i = 0; i < 8; i++) memcpy(b + 4 * i, &r.v[i], 4); }
extern "C" {
// all I/O: 8 little-endian u32 limbs
void hs_fe_mul(const uint8_t* a, const uint8_t* b, uint8_t* r) { fe x, y, z; load(x, a); load(y, b); fe_mul(z, x, y); store(r, z); }
void hs_fe_sqr(const uint8_t* a, uint8_t* r) { fe x, z; load(x, a); fe_sqr(z, x); store(r, z); }
void hs_fe_add(const uint8_t* a, const uint8_t* b, uint8_t* r) { fe x, y, z; load(x, a); load(y, b); fe_add(z, x, y); store(r, z); }
void hs_fe_sub(const uint8_t* a, const uint8_t* b, uint8_t* r) { fe x, y, z; load(x, a); load(y, b); fe_sub(z, x, y); store(r, z); }
void hs_fe_mul8(const uint8_t* a, uint8_t* r) { fe x, z; load(x, a); fe_mul8(z, x); store(r, z); }
void hs_fe_mul3(const uint8_t* a, uint8_t* r) { fe x, z; load(x, a); fe_mul3(z, x); store(r, z); }
void hs_fe_norm(const uint8_t* a, uint8_t* r) { fe x; load(x, a); fe_normalize(x); store(r, x); }
void hs_fe_inv(const uint8_t* a, uint8_t* r) { fe x, z; load(x, a); fe_inv(z, x); store(r, z); }
int hs_fe_sqrt(const uint8_t* a, uint8_t* r) { fe x, z; load(x, a); bool ok = fe_sqrt(z, x); store(r, z); return ok; }
int hs_fe_is_zero(const uint8_t* a) { fe x; load(x, a); return fe_is_zero(x); }
void hs_mul_wide(const uint8_t* a, const uint8_t* b, uint8_t* r) { fe x, y; load(x, a); load(y, b); uint32_t t[16]; mul_wide(t, x.v, y.v); memcpy(r, t, 64); }
void hs_fe_reduce_wide(const uint8_t* t64, uint8_t* r) { uint32_t t[16]; memcpy(t, t64, 64); fe z; fe_reduce_wide(z, t); store(r, z); }
void hs_sqr_wide(const uint8_t* a, uint8_t* r) { fe x; load(x, a); uint32_t t[16]; sqr_wide(t, x.v); memcpy(r, t, 64); }
}
