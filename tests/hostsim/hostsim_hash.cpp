// Host build of the device hashing headers for GPU-less unit tests (TEST BUILD ONLY).
#include "../../rusty_kaspa_b200/csrc/kgv_txhash.cuh"
#include <cstring>
using namespace kgv;
extern "C" {
void hs_b2b(uint32_t domain, const uint8_t* data, uint32_t n, uint8_t* out) {
  Blake2b s; b2b_init(s, domain); b2b_bytes(s, data, n);
  uint64_t d[4]; b2b_final(s, d); memcpy(out, d, 32);
}
void hs_tx_id(const kgv_tx_batch* b, uint32_t tx, uint8_t* out) {
  BatchView v{b->txs, b->inputs, b->outputs, b->entries, b->bytes};
  uint64_t d[4]; tx_id(d, v, tx); memcpy(out, d, 32);
}
void hs_tx_hash(const kgv_tx_batch* b, uint32_t tx, uint8_t* out) {
  BatchView v{b->txs, b->inputs, b->outputs, b->entries, b->bytes};
  uint64_t d[4]; tx_hash(d, v, tx); memcpy(out, d, 32);
}
void hs_sighash(const kgv_tx_batch* b, uint32_t tx, uint32_t in_abs, uint32_t hash_type, int ecdsa, uint8_t* out) {
  BatchView v{b->txs, b->inputs, b->outputs, b->entries, b->bytes};
  SigHashReused r; sighash_reused(r, v, tx);
  uint32_t w[8]; sighash_final(w, v, tx, in_abs, hash_type, ecdsa != 0, r);
  for (int i = 0; i < 8; i++) { out[4 * i] = w[i] >> 24; out[4 * i + 1] = w[i] >> 16; out[4 * i + 2] = w[i] >> 8; out[4 * i + 3] = w[i]; }
}
}
