// Host build of the MuHash field arithmetic / element expansion (rusty_kaspa_b200/csrc/kgv_u3072.cuh) for GPU-less
// unit tests (tests/test_hostsim.py).  Test infrastructure only.
#include <string.h>
#include "../../rusty_kaspa_b200/csrc/kgv_u3072.cuh"
#include "../../rusty_kaspa_b200/csrc/kgv_muhash.cuh"
using namespace kgv;
extern "C" {
// contiguous 96-limb little-endian numbers (stride 1)
void hs_u3072_mul_mod(const uint32_t* a, const uint32_t* b, uint32_t* r, uint32_t* wide192) {
  uint32_t P[192];
  u3072_mul_mod(r, 1, 0, P, 1, 0, a, 1, 0, b, 1, 0);
  if (wide192) memcpy(wide192, P, sizeof P);
}
void hs_u3072_fold(const uint32_t* wide192, uint32_t* r) { u3072_fold(r, 1, 0, wide192, 1, 0); }
void hs_u3072_canonical(const uint32_t* a, uint32_t* out) { u3072_canonical(out, a, 1, 0); }
// strided layout check: element e of an array with `stride` elements
void hs_u3072_mul_mod_strided(uint32_t* arr, size_t stride, size_t ea, size_t eb, size_t er, uint32_t* scratch, size_t sstride, size_t se) {
  u3072_mul_mod(arr, stride, er, scratch, sstride, se, arr, stride, ea, arr, stride, eb);
}
void hs_muhash_expand(const uint8_t* digest32, uint32_t* out96) {
  uint64_t d[4];
  memcpy(d, digest32, 32);
  muhash_expand_store(out96, 1, 0, d);
}
}
extern "C" {
// the cooperative multiplier with its 16 lanes run in a loop, phase by phase
int hs_u3072_coop_mul_mod(const uint32_t* a, const uint32_t* b, uint32_t* r) {
  static U3072Coop s;
  for (int i = 0; i < 96; i++) { s.a[i] = a[i]; s.b[i] = b[i]; }
  for (int l = 0; l < 16; l++) u3072_coop_phase1(l, s);
  for (int l = 0; l < 16; l++) u3072_coop_phase2(l, s);
  int bad = s.cols[22][16] != 0;
  for (int l = 0; l < 16; l++) u3072_coop_phase2b(l, s);
  for (int l = 0; l < 16; l++) u3072_coop_phase3(l, s);
  for (int l = 0; l < 16; l++) u3072_coop_phase3b(l, s);
  for (int i = 0; i < 96; i++) r[i] = s.w[i / 8][i % 8];
  return bad;
}
}
extern "C" {
// keyed BLAKE2b of the two MuHash domains (which = 0 "MuHashElement", 1 "MuHashFinalize") over raw bytes
void hs_muhash_domain_hash(int which, const uint8_t* data, size_t n, uint8_t* out32) {
  Blake2b h;
  if (which) b2b_init_muhash_finalize(h); else b2b_init_muhash_element(h);
  for (size_t i = 0; i < n; i++) b2b_byte(h, data[i]);
  uint64_t d[4];
  b2b_final(h, d);
  memcpy(out32, d, 32);
}
}
