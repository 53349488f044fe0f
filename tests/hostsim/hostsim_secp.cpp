// Host build of the device verification cores (portable primitive bodies) for GPU-less unit tests.
// TEST BUILD ONLY: never linked into the product library.
#include "../../rusty_kaspa_b200/csrc/kgv_verify.cuh"
#include <cstring>
#include <vector>
using namespace kgv;

struct HostTab {
  uint32_t d[8][16];
  void put(int e, int w, uint32_t v) { d[e][w] = v; }
  uint32_t get(int e, int w) const { return d[e][w]; }
};
// generator-table entries are computed on demand (the device reads a prebuilt table)
struct HostGLoad {
  void operator()(fe& x, fe& y, const uint32_t* entry) const {
    size_t idx = (size_t)(entry - (const uint32_t*)nullptr) / 16;
    const fe g = {KGV_GX_LIMBS}, gy = {KGV_GY_LIMBS}, h = {KGV_G128X_LIMBS}, hy = {KGV_G128Y_LIMBS};
    if (idx < 65536) gtab_entry(x, y, (uint32_t)idx, g, gy);
    else gtab_entry(x, y, (uint32_t)(idx - 65536), h, hy);
  }
};
static void be_words(uint32_t* w, const uint8_t* b, int n) {
  for (int i = 0; i < n; i++) w[i] = ((uint32_t)b[4 * i] << 24) | ((uint32_t)b[4 * i + 1] << 16) | ((uint32_t)b[4 * i + 2] << 8) | b[4 * i + 3];
}
extern "C" {
int hs_schnorr_verify(const uint8_t* pk, const uint8_t* msg, const uint8_t* sig) {
  uint32_t pkw[8], mw[8], sw[16];
  be_words(pkw, pk, 8); be_words(mw, msg, 8); be_words(sw, sig, 16);
  HostTab tab;
  return schnorr_verify_core(pkw, mw, sw, tab, (const uint32_t*)nullptr, HostGLoad());
}
struct HostTrace {
  uint32_t* out;
  void operator()(int stage, const uint32_t* w, int n) const { for (int i = 0; i < n && i < 16; i++) out[stage * 16 + i] = w[i]; }
};
// same layout as kgv_debug_schnorr_trace: out[32][16]
int hs_schnorr_trace(const uint8_t* pk, const uint8_t* msg, const uint8_t* sig, uint32_t* out) {
  uint32_t pkw[8], mw[8], sw[16];
  be_words(pkw, pk, 8); be_words(mw, msg, 8); be_words(sw, sig, 16);
  HostTab tab;
  memset(out, 0, 32 * 16 * 4);
  return schnorr_verify_core(pkw, mw, sw, tab, (const uint32_t*)nullptr, HostGLoad(), HostTrace{out});
}
int hs_ecdsa_verify(const uint8_t* pk33, const uint8_t* msg, const uint8_t* sig) {
  uint32_t pkw[8], mw[8], sw[16];
  be_words(pkw, pk33 + 1, 8); be_words(mw, msg, 8); be_words(sw, sig, 16);
  HostTab tab;
  return ecdsa_verify_core(pk33[0], pkw, mw, sw, tab, (const uint32_t*)nullptr, HostGLoad());
}
// k: 8 LE limbs -> k1[5], k2[5] LE limbs, signs
void hs_glv_split(const uint32_t* k, uint32_t* k1, int* n1, uint32_t* k2, int* n2) {
  bool a, b;
  glv_split(k1, a, k2, b, k);
  *n1 = a; *n2 = b;
}
void hs_sc_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { sc_mul(r, a, b); }
void hs_sc_inv(const uint32_t* a, uint32_t* r) { sc_inv(r, a); }
void hs_gtab_entry(uint32_t v, int which, uint32_t* xy) {
  const fe g = {KGV_GX_LIMBS}, gy = {KGV_GY_LIMBS}, h = {KGV_G128X_LIMBS}, hy = {KGV_G128Y_LIMBS};
  fe x, y;
  if (which == 0) gtab_entry(x, y, v, g, gy); else gtab_entry(x, y, v, h, hy);
  memcpy(xy, x.v, 32); memcpy(xy + 8, y.v, 32);
}
void hs_sha_challenge(const uint8_t* r, const uint8_t* pk, const uint8_t* m, uint8_t* out) {
  uint32_t rw[8], pw[8], mw[8], o[8];
  be_words(rw, r, 8); be_words(pw, pk, 8); be_words(mw, m, 8);
  bip340_challenge(o, rw, pw, mw);
  for (int i = 0; i < 8; i++) { out[4 * i] = o[i] >> 24; out[4 * i + 1] = o[i] >> 16; out[4 * i + 2] = o[i] >> 8; out[4 * i + 3] = o[i]; }
}
void hs_ecdsa_wrap(const uint8_t* h, uint8_t* out) {
  uint32_t hw[8], o[8];
  be_words(hw, h, 8);
  ecdsa_sighash_wrap(o, hw);
  for (int i = 0; i < 8; i++) { out[4 * i] = o[i] >> 24; out[4 * i + 1] = o[i] >> 16; out[4 * i + 2] = o[i] >> 8; out[4 * i + 3] = o[i]; }
}
}
