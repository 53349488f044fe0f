// kgv_context.cuh — the UTXO-context rules of one transaction (K6), shared by the batch kernel (k_tx_context) and the
// in-order replay kernel.  Follows validate_populated_transaction_and_get_fee
// (consensus/src/processes/transaction_validator/tx_validation_in_utxo_context.rs:34-61): coinbase maturity :75-91,
// input amounts :93-108, output amounts / fee :110-118, storage mass :120-128, sequence locks :130-155; the populate step
// that precedes it (first missing entry => MissingTxOutpoints) is utxo_validation.rs:319-327.
#pragma once
#include "kgv_txhash.cuh"

namespace kgv {

// b.entries must hold one DevEntry per input.  `skip` marks the transaction as a coinbase to be skipped
// (utxo_validation.rs:273 skips position 0; the batch kernels recognise it by its subnetwork id).
__device__ __forceinline__ kgv_tx_result tx_context_rules(const BatchView& b, uint32_t ti, uint64_t pov, uint32_t flags, const kgv_params& prm, bool skip) {
  const kgv_tx& t = b.txs[ti];
  kgv_tx_result r;
  r.fee = 0; r.fail_input = 0; r.status = KGV_TX_OK; r.script_err = 0; r.pad_[0] = r.pad_[1] = 0;
  const DevEntry* ent = b.entries + t.first_input;
  if (skip) { r.status = KGV_TX_SKIPPED_COINBASE; return r; }
  for (uint32_t i = 0; i < t.n_inputs; i++)
    if (!ent[i].found) { r.status = KGV_TX_MISSING_OUTPOINTS; return r; }  // utxo_validation.rs:319-327
  if (flags == KGV_FLAGS_SCRIPTS_ONLY) return r;
  for (uint32_t i = 0; i < t.n_inputs; i++)
    if (ent[i].is_coinbase && ent[i].block_daa_score + prm.coinbase_maturity > pov) { r.status = KGV_TX_IMMATURE_COINBASE; r.fail_input = i; return r; }
  uint64_t total_in = 0;
  for (uint32_t i = 0; i < t.n_inputs; i++) {
    if (ck_add(total_in, ent[i].amount, total_in)) { r.status = KGV_TX_INPUT_AMOUNT_OVERFLOW; return r; }
    if (total_in > prm.max_sompi) { r.status = KGV_TX_INPUT_AMOUNT_TOO_HIGH; return r; }
  }
  uint64_t total_out = 0;
  for (uint32_t i = 0; i < t.n_outputs; i++) total_out += b.outputs[t.first_output + i].value;
  if (total_in < total_out) { r.status = KGV_TX_SPEND_TOO_HIGH; return r; }
  r.fee = total_in - total_out;
  if (flags != KGV_FLAGS_SKIP_MASS_CHECK) {
    uint64_t mass;
    const kgv_output* outs = b.outputs + t.first_output;
    bool ok = storage_mass(mass, false, t.n_inputs, t.n_outputs, [&](uint32_t i) -> const DevEntry& { return ent[i]; },
                           [&](uint32_t i, uint64_t& v, uint32_t& l) { v = outs[i].value; l = outs[i].script_len; }, prm.storage_mass_parameter);
    if (!ok) { r.status = KGV_TX_MASS_INCOMPUTABLE; return r; }
    if (mass != t.mass) { r.status = KGV_TX_WRONG_MASS; return r; }
  }
  for (uint32_t i = 0; i < t.n_inputs; i++) {
    uint64_t seq = b.inputs[t.first_input + i].sequence;
    if (seq & (1ull << 63)) continue;
    long long lock = (long long)ent[i].block_daa_score + (long long)(seq & 0xFFFFFFFFull) - 1;
    if (lock >= (long long)pov) { r.status = KGV_TX_SEQUENCE_LOCK; return r; }
  }
  return r;
}

}  // namespace kgv
