// Host build of the device hashing headers for GPU-less unit tests (TEST BUILD ONLY).
#include "../../rusty_kaspa_b200/csrc/kgv_txhash.cuh"
#include <cstring>
#include <vector>
using namespace kgv;
extern "C" {
void hs_b2b(uint32_t domain, const uint8_t* data, uint32_t n, uint8_t* out) {
  Blake2b s; b2b_init(s, domain); b2b_bytes(s, data, n);
  uint64_t d[4]; b2b_final(s, d); memcpy(out, d, 32);
}
static std::vector<DevEntry> dev_entries(const kgv_tx_batch* b) {
  std::vector<DevEntry> v(b->n_inputs);
  for (size_t i = 0; b->entries && i < b->n_inputs; i++) {
    const kgv_utxo_entry& e = b->entries[i];
    v[i] = DevEntry{e.amount, e.block_daa_score, b->bytes + e.script_off, e.script_len, e.spk_version, e.is_coinbase, 1};
  }
  return v;
}
void hs_tx_id(const kgv_tx_batch* b, uint32_t tx, uint8_t* out) {
  BatchView v{b->txs, b->inputs, b->outputs, nullptr, b->bytes};
  uint64_t d[4]; tx_id(d, v, tx); memcpy(out, d, 32);
}
void hs_tx_hash(const kgv_tx_batch* b, uint32_t tx, uint8_t* out) {
  BatchView v{b->txs, b->inputs, b->outputs, nullptr, b->bytes};
  uint64_t d[4]; tx_hash(d, v, tx); memcpy(out, d, 32);
}
void hs_sighash(const kgv_tx_batch* b, uint32_t tx, uint32_t in_abs, uint32_t hash_type, int ecdsa, uint8_t* out) {
  std::vector<DevEntry> de = dev_entries(b);
  BatchView v{b->txs, b->inputs, b->outputs, de.data(), b->bytes};
  SigHashReused r; sighash_reused(r, v, tx);
  uint32_t w[8]; sighash_final(w, v, tx, in_abs, hash_type, ecdsa != 0, r);
  for (int i = 0; i < 8; i++) { out[4 * i] = w[i] >> 24; out[4 * i + 1] = w[i] >> 16; out[4 * i + 2] = w[i] >> 8; out[4 * i + 3] = w[i]; }
}

// standard-class script check of one input, CPU-simulating plan -> (sighash + verify via callbacks) -> resolve.
// verdicts for the candidate pairs are supplied by the caller (computed with the C oracle in the test).
int hs_plan_input(const kgv_tx_batch* b, uint32_t in_abs, uint8_t* plan_out /*cls,m,n,n_items*/) {
  std::vector<DevEntry> de = dev_entries(b);
  const kgv_input& in = b->inputs[in_abs];
  InputPlan pl;
  plan_input(pl, b->bytes + in.sigscript_off, in.sigscript_len, de[in_abs]);
  plan_out[0] = pl.cls; plan_out[1] = pl.m; plan_out[2] = pl.n; plan_out[3] = pl.n_items;
  return pl.cls;
}
// item k of an input -> copies sig(65 incl. hash type) and key (32/33) out
void hs_item(const kgv_tx_batch* b, uint32_t in_abs, uint32_t k, uint8_t* sig65, uint8_t* key33) {
  std::vector<DevEntry> de = dev_entries(b);
  const kgv_input& in = b->inputs[in_abs];
  InputPlan pl;
  plan_input(pl, b->bytes + in.sigscript_off, in.sigscript_len, de[in_abs]);
  const uint8_t *s, *kk;
  item_location(pl, k, b->bytes + in.sigscript_off, de[in_abs], s, kk);
  memcpy(sig65, s, 65);
  memcpy(key33, kk, (pl.cls == CLS_P2PK_ECDSA || pl.cls == CLS_MULTISIG_ECDSA) ? 33 : 32);
}
uint32_t hs_resolve_input(const kgv_tx_batch* b, uint32_t in_abs, const uint8_t* verdicts) {
  std::vector<DevEntry> de = dev_entries(b);
  const kgv_input& in = b->inputs[in_abs];
  InputPlan pl;
  plan_input(pl, b->bytes + in.sigscript_off, in.sigscript_len, de[in_abs]);
  return resolve_input(pl, b->bytes + in.sigscript_off, de[in_abs], in.sig_op_count, verdicts);
}
}
