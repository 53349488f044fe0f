// kgv_txhash.cuh — transaction id / transaction hash / signature hash, one hash per thread.
//
// Byte-exact restatement for the GPU of
//   consensus/core/src/hashing/tx.rs:16-107        tx id (EXCLUDE_SIGNATURE_SCRIPT|EXCLUDE_MASS_COMMIT
//                                                  unless coinbase) and tx hash (FULL)
//   consensus/core/src/hashing/sighash.rs:140-277  the five reusable sub-hashes and the final
//                                                  keyed BLAKE2b; ECDSA adds the SHA-256 wrap
//   consensus/core/src/hashing/mod.rs:46-96        lengths as u64 LE, var-bytes = len || bytes
// operating directly on the flat SoA batch of include/kgv.h.
#pragma once
#include "../../include/kgv.h"
#include "kgv_blake2b.cuh"
#include "kgv_sha256.cuh"
#include "kgv_script_std.cuh"

namespace kgv {

struct BatchView {
  const kgv_tx* txs;
  const kgv_input* inputs;
  const kgv_output* outputs;
  const DevEntry* entries;  // populated entry per input (null when not needed)
  const uint8_t* bytes;
};

KGV_HD bool tx_is_coinbase(const kgv_tx& t) {  // subnets::SUBNETWORK_ID_COINBASE = 01 00 .. 00
  bool r = t.subnetwork_id[0] == 1;
#pragma unroll
  for (int i = 1; i < 20; i++) r = r && t.subnetwork_id[i] == 0;
  return r;
}
KGV_HD bool tx_is_native(const kgv_tx& t) {
  bool r = true;
#pragma unroll
  for (int i = 0; i < 20; i++) r = r && t.subnetwork_id[i] == 0;
  return r;
}
KGV_HD void b2b_var_bytes(Blake2b& s, const uint8_t* p, uint32_t n) {
  b2b_u64(s, n);
  b2b_bytes(s, p, n);
}
KGV_HD void hash_output(Blake2b& s, const BatchView& b, const kgv_output& o) {
  b2b_u64(s, o.value);
  b2b_u16(s, o.spk_version);
  b2b_var_bytes(s, b.bytes + o.script_off, o.script_len);
}

// hashing/tx.rs:45-107
KGV_HD void write_transaction(Blake2b& s, const BatchView& b, const kgv_tx& t, bool exclude_sigscript, bool exclude_mass) {
  b2b_u16(s, t.version);
  b2b_u64(s, t.n_inputs);
  for (uint32_t i = 0; i < t.n_inputs; i++) {
    const kgv_input& in = b.inputs[t.first_input + i];
    b2b_bytes(s, in.prev_txid, 32);
    b2b_u32(s, in.prev_index);
    if (!exclude_sigscript) {
      b2b_var_bytes(s, b.bytes + in.sigscript_off, in.sigscript_len);
      b2b_u8(s, in.sig_op_count);
    } else {
      b2b_u64(s, 0);
    }
    b2b_u64(s, in.sequence);
  }
  b2b_u64(s, t.n_outputs);
  for (uint32_t i = 0; i < t.n_outputs; i++) hash_output(s, b, b.outputs[t.first_output + i]);
  b2b_u64(s, t.lock_time);
  b2b_bytes(s, t.subnetwork_id, 20);
  b2b_u64(s, t.gas);
  b2b_var_bytes(s, b.bytes + t.payload_off, t.payload_len);
  if (!exclude_mass && t.mass > 0) b2b_u64(s, t.mass);
}

// out: 4 LE u64 words = 32 bytes
KGV_HD void tx_id(uint64_t* out, const BatchView& b, uint32_t tx) {
  const kgv_tx& t = b.txs[tx];
  bool cb = tx_is_coinbase(t);
  Blake2b s;
  b2b_init(s, B2B_TX_ID);
  write_transaction(s, b, t, !cb, !cb);
  b2b_final(s, out);
}
KGV_HD void tx_hash(uint64_t* out, const BatchView& b, uint32_t tx) {
  Blake2b s;
  b2b_init(s, B2B_TX_HASH);
  write_transaction(s, b, b.txs[tx], false, false);
  b2b_final(s, out);
}

// The per-transaction reusable values (sighash.rs:14-138 SigHashReusedValues): computed once per tx.
struct SigHashReused {
  uint64_t prevouts[4], sequences[4], sigopcounts[4], outputs[4], payload[4];
};
KGV_HD void sighash_reused(SigHashReused& r, const BatchView& b, uint32_t tx) {
  const kgv_tx& t = b.txs[tx];
  Blake2b s;
  b2b_init(s, B2B_SIGHASH);  // sighash.rs:140-153
  for (uint32_t i = 0; i < t.n_inputs; i++) { const kgv_input& in = b.inputs[t.first_input + i]; b2b_bytes(s, in.prev_txid, 32); b2b_u32(s, in.prev_index); }
  b2b_final(s, r.prevouts);
  b2b_init(s, B2B_SIGHASH);  // :155-167
  for (uint32_t i = 0; i < t.n_inputs; i++) b2b_u64(s, b.inputs[t.first_input + i].sequence);
  b2b_final(s, r.sequences);
  b2b_init(s, B2B_SIGHASH);  // :169-182
  for (uint32_t i = 0; i < t.n_inputs; i++) b2b_u8(s, b.inputs[t.first_input + i].sig_op_count);
  b2b_final(s, r.sigopcounts);
  b2b_init(s, B2B_SIGHASH);  // :197-221 (hash of all outputs)
  for (uint32_t i = 0; i < t.n_outputs; i++) hash_output(s, b, b.outputs[t.first_output + i]);
  b2b_final(s, r.outputs);
  if (tx_is_native(t) && t.payload_len == 0) {  // :184-195
#pragma unroll
    for (int i = 0; i < 4; i++) r.payload[i] = 0;
  } else {
    b2b_init(s, B2B_SIGHASH);
    b2b_var_bytes(s, b.bytes + t.payload_off, t.payload_len);
    b2b_final(s, r.payload);
  }
}

// Final signature hash of one input (sighash.rs:238-277).  `in_abs` indexes b.inputs / b.entries,
// hash_type must be one of the six allowed values (callers check, sighash_type.rs:15-22,50-56).
// out: 8 big-endian numeric words (the form the verification cores consume).
KGV_HD void sighash_final(uint32_t* out_be_words, const BatchView& b, uint32_t tx, uint32_t in_abs, uint32_t hash_type, bool ecdsa,
                          const SigHashReused& r) {
  const kgv_tx& t = b.txs[tx];
  const kgv_input& in = b.inputs[in_abs];
  const DevEntry& e = b.entries[in_abs];
  const bool acp = (hash_type & 0x80u) != 0;
  const uint32_t base = hash_type & 7u;
  const uint64_t zero[4] = {0, 0, 0, 0};
  uint64_t single_out[4] = {0, 0, 0, 0};
  const uint64_t* outs = r.outputs;
  if (base == 2u) {
    outs = zero;
  } else if (base == 4u) {
    uint32_t rel = in_abs - t.first_input;
    if (rel < t.n_outputs) {
      Blake2b so;
      b2b_init(so, B2B_SIGHASH);
      hash_output(so, b, b.outputs[t.first_output + rel]);
      b2b_final(so, single_out);
    }
    outs = single_out;
  }
  Blake2b s;
  b2b_init(s, B2B_SIGHASH);
  b2b_u16(s, t.version);
  b2b_digest_words(s, acp ? zero : r.prevouts);
  b2b_digest_words(s, (acp || base == 4u || base == 2u) ? zero : r.sequences);
  b2b_digest_words(s, acp ? zero : r.sigopcounts);
  b2b_bytes(s, in.prev_txid, 32);
  b2b_u32(s, in.prev_index);
  b2b_u16(s, e.spk_version);
  b2b_var_bytes(s, e.script, e.script_len);
  b2b_u64(s, e.amount);
  b2b_u64(s, in.sequence);
  b2b_u8(s, in.sig_op_count);
  b2b_digest_words(s, outs);
  b2b_u64(s, t.lock_time);
  b2b_bytes(s, t.subnetwork_id, 20);
  b2b_u64(s, t.gas);
  b2b_digest_words(s, r.payload);
  b2b_u8(s, hash_type);
  uint64_t d[4];
  b2b_final(s, d);
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 4; i++) {  // digest bytes -> big-endian numeric words
    w[2 * i] = bswap32((uint32_t)d[i]);
    w[2 * i + 1] = bswap32((uint32_t)(d[i] >> 32));
  }
  if (ecdsa) ecdsa_sighash_wrap(out_be_words, w);
  else {
#pragma unroll
    for (int i = 0; i < 8; i++) out_be_words[i] = w[i];
  }
}

}  // namespace kgv
