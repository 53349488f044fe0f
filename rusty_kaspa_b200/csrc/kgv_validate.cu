// kgv_validate.cu — GPU-resident UTXO table, per-transaction context rules and the fused
// validate_transactions path (include/kgv.h).
//
// Data flow of kgv_validate_txs (one call = one block's worth or more of transactions):
//   k_populate        per input : outpoint -> table slot -> DevEntry            (HBM latency bound)
//   k_tx_context      per tx    : maturity, amounts, storage mass, seq-lock      (tx_validation_in_utxo_context.rs:75-155)
//   k_plan            per input : recognise P2PK / P2PK-ECDSA / P2SH multisig, count signature checks
//   (scan)                        exclusive prefix sums -> item offsets (two lists: Schnorr, ECDSA)
//   k_emit_items      per input : gather (pk, sig) of every candidate pair into SoA arrays
//   k_sighash_reused  per tx    : the five reusable sub-hashes                   (sighash.rs:14-221)
//   k_item_msgs       per item  : final signature hash                           (sighash.rs:238-277)
//   k_schnorr_verify / k_ecdsa_verify per item                                   (lib.rs:593, :628)
//   k_resolve         per input : replay of the script engine over the verdicts  (lib.rs:488-571)
//   k_tx_finalize     per tx    : first failing input -> TxRuleError class       (:162-200)
#include "kgv_internal.h"
#include "kgv_muhash.cuh"
#include "kgv_txhash.cuh"
#include "kgv_utxo.cuh"
#include "kgv_context.cuh"

#include <cstdio>
#include <vector>

using namespace kgv;

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      char b_[256];                                                                               \
      snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->err = b_;                                                                              \
      return KGV_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// KGV_DEBUG=1: synchronise and report after every stage of the fused path (locates a faulting kernel)
#include <cstdlib>
static bool kgv_debug_on() { static int v = -1; if (v < 0) v = getenv("KGV_DEBUG") ? 1 : 0; return v == 1; }
#define STAGE(name)                                                                               \
  do {                                                                                            \
    if (kgv_debug_on()) {                                                                         \
      cudaError_t e_ = cudaStreamSynchronize(ctx->stream);                                        \
      fprintf(stderr, "[kgv] stage %s: %s\n", name, cudaGetErrorString(e_));                      \
      fflush(stderr);                                                                             \
    }                                                                                             \
  } while (0)

// ---------------------------------------------------------------------------------------------
// UTXO table kernels (device functions: kgv_utxo.cuh)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void slot_to_entry(DevEntry& e, const TableView& t, const UtxoSlot* s) {
  SlotHead h;
  slot_load_head(h, s);
  head_to_entry(e, t, s, h);
}

__global__ void k_utxo_lookup(TableView t, const uint8_t* __restrict__ keys, size_t n, kgv_utxo_entry* __restrict__ entries,
                              uint8_t* __restrict__ scripts, uint32_t stride, uint8_t* __restrict__ found) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[9];
  load_key(k, keys + 36 * i);
  SlotHead h;
  UtxoSlot* s = table_find(t, k, h);
  // the 32-byte entry record is written with one 256-bit store
  uint32_t e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  e[4] = (uint32_t)(i * stride);  // script_off
  if (s) {
    DevEntry d;
    head_to_entry(d, t, s, h);
    e[0] = h.w[10]; e[1] = h.w[11]; e[2] = h.w[12]; e[3] = h.w[13];
    e[5] = d.script_len;
    e[6] = (uint32_t)d.spk_version | ((uint32_t)d.is_coinbase << 16);
    for (uint32_t b = 0; b < d.script_len && b < stride; b++) scripts[i * stride + b] = d.script[b];
  }
  if ((((uintptr_t)entries) & 31) == 0) st256(entries + i, e);
  else {
    uint32_t* o = (uint32_t*)(entries + i);
#pragma unroll
    for (int j = 0; j < 8; j++) o[j] = e[j];
  }
  found[i] = s ? 1 : 0;
}
__global__ void k_utxo_erase(TableView t, const uint8_t* __restrict__ keys, size_t n, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[9];
  load_key(k, keys + 36 * i);
  uint32_t r = table_erase(t, k);
  if (status) status[i] = (uint8_t)r;
}
__global__ void k_utxo_insert(TableView t, const uint8_t* __restrict__ keys, const kgv_utxo_entry* __restrict__ entries, const uint8_t* __restrict__ bytes,
                              size_t n, uint8_t* __restrict__ status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[9];
  load_key(k, keys + 36 * i);
  kgv_utxo_entry e = entries[i];
  uint32_t r = table_put(t, k, e.amount, e.block_daa_score, e.spk_version, e.is_coinbase, bytes + e.script_off, e.script_len);
  if (status) status[i] = (uint8_t)r;
}

// export (DbUtxoSetStore::iterator, utxo_set.rs:114-129: the syncer side of a pruning-point import and the source of `virtual.utxo_set := pruning
// utxo_set`, processor.rs:1150-1158): every live slot is compacted into (key, entry, script bytes) arrays; order is the table's, i.e. arbitrary
__global__ void k_utxo_export(TableView t, uint8_t* __restrict__ keys, kgv_utxo_entry* __restrict__ entries, uint8_t* __restrict__ bytes, uint64_t max_n, uint64_t bytes_cap,
                              unsigned long long* __restrict__ cnt) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > t.mask) return;
  const UtxoSlot* s = &t.slots[i];
  if (s->state != SLOT_FULL) return;
  DevEntry e;
  slot_to_entry(e, t, s);
  const unsigned long long at = atomicAdd(&cnt[0], 1ull);
  const unsigned long long off = atomicAdd(&cnt[1], (unsigned long long)e.script_len);
  if (!keys || at >= max_n || off + e.script_len > bytes_cap || off + e.script_len > 0xFFFFFFFFull) return;  // counting pass / caller's arrays too small (reported by the host side)
  uint32_t* kw = (uint32_t*)(keys + 36 * at);
#pragma unroll
  for (int w = 0; w < 9; w++) kw[w] = s->key[w];
  kgv_utxo_entry o;
  o.amount = e.amount; o.block_daa_score = e.block_daa_score; o.script_off = (uint32_t)off; o.script_len = e.script_len; o.spk_version = e.spk_version; o.is_coinbase = e.is_coinbase;
  memset(o.pad_, 0, sizeof o.pad_);
  entries[at] = o;
  for (uint32_t b = 0; b < e.script_len; b++) bytes[off + b] = e.script[b];
}
// MuHash::from_utxo of every (outpoint, entry) of a chunk (consensus/src/consensus/mod.rs:1075-1080): level 0 of a product tree
__global__ void __launch_bounds__(128) k_muhash_utxo_elements(const uint8_t* __restrict__ keys, const kgv_utxo_entry* __restrict__ entries, const uint8_t* __restrict__ bytes, size_t n,
                                                              uint32_t* __restrict__ e_num) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  uint32_t k[9];
  load_key(k, keys + 36 * g);
  const kgv_utxo_entry e = entries[g];
  uint64_t d[4];
  muhash_utxo_digest(d, k, k[8], e.block_daa_score, e.amount, e.is_coinbase != 0, e.spk_version, bytes + e.script_off, e.script_len);
  muhash_expand_store(e_num, n, g, d);
}

// digest: sum of MuHashElement hashes, accumulated as 8 x 32-bit limbs in 64-bit counters (carries folded on the host)
__global__ void k_utxo_digest(TableView t, unsigned long long* __restrict__ acc) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > t.mask) return;
  const UtxoSlot* s = &t.slots[i];
  if (s->state != SLOT_FULL) return;
  DevEntry e;
  slot_to_entry(e, t, s);
  Blake2b h;
  b2b_init_muhash_element(h);
  for (int w = 0; w < 9; w++) b2b_u32(h, s->key[w]);
  b2b_u64(h, e.block_daa_score);
  b2b_u64(h, e.amount);
  b2b_u8(h, e.is_coinbase ? 1 : 0);
  b2b_u16(h, e.spk_version);
  b2b_u64(h, e.script_len);
  b2b_bytes(h, e.script, e.script_len);
  uint64_t d[4];
  b2b_final(h, d);
#pragma unroll
  for (int w = 0; w < 4; w++) {
    atomicAdd(&acc[2 * w], (unsigned long long)(uint32_t)d[w]);
    atomicAdd(&acc[2 * w + 1], (unsigned long long)(uint32_t)(d[w] >> 32));
  }
}

// ---------------------------------------------------------------------------------------------
// validation kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_entries_from_batch(const kgv_utxo_entry* __restrict__ in, const uint8_t* __restrict__ bytes, size_t n, DevEntry* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  kgv_utxo_entry e = in[i];
  DevEntry d;
  d.amount = e.amount; d.block_daa_score = e.block_daa_score; d.script = bytes + e.script_off; d.script_len = e.script_len;
  d.spk_version = e.spk_version; d.is_coinbase = e.is_coinbase; d.found = e.pad_[0] ? 0 : 1;  // pad_[0] != 0: caller marks the entry absent
  out[i] = d;
}

__global__ void k_populate(TableView t, const kgv_input* __restrict__ inputs, size_t n, DevEntry* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[9];
  input_key(k, inputs[i]);
  SlotHead h;
  UtxoSlot* s = table_find(t, k, h);
  DevEntry d;
  if (s) head_to_entry(d, t, s, h);
  else entry_absent(d);
  out[i] = d;
}

__global__ void k_tx_context(BatchView b, uint32_t n_txs, uint64_t pov, uint32_t flags, kgv_params prm, kgv_tx_result* __restrict__ res) {
  uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
  if (ti >= n_txs) return;
  res[ti] = tx_context_rules(b, ti, pov, flags, prm, tx_is_coinbase(b.txs[ti]));
}

// plan: one thread per input. counts[0][i] = Schnorr items, counts[1][i] = ECDSA items (0 when the tx already failed).
__global__ void k_plan(BatchView b, size_t n_inputs, const uint32_t* __restrict__ input_tx, const kgv_tx_result* __restrict__ res,
                       InputPlan* __restrict__ plans, uint32_t* __restrict__ cnt_s, uint32_t* __restrict__ cnt_e) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inputs) return;
  InputPlan pl;
  pl.item_base = 0; pl.redeem_off = 0; pl.redeem_len = 0; pl.cls = CLS_NONSTANDARD; pl.m = pl.n = pl.n_items = 0; pl.pad_[0] = pl.pad_[1] = 0;
  uint32_t cs = 0, ce = 0;
  if (res[input_tx[i]].status == KGV_TX_OK) {
    const kgv_input& in = b.inputs[i];
    plan_input(pl, b.bytes + in.sigscript_off, in.sigscript_len, b.entries[i]);
    if (pl.cls == CLS_P2PK || pl.cls == CLS_MULTISIG) cs = pl.n_items;
    if (pl.cls == CLS_P2PK_ECDSA || pl.cls == CLS_MULTISIG_ECDSA) ce = pl.n_items;
  }
  plans[i] = pl;
  cnt_s[i] = cs;
  cnt_e[i] = ce;
}

// exclusive scans of the two per-input item-count arrays (Schnorr / ECDSA), one block each in ONE launch: every thread sums a
// contiguous chunk, the 1024 chunk sums are scanned in shared memory, then each thread rewrites its chunk.  (Plumbing: a few
// microseconds for a few hundred thousand inputs; the first version scanned 1024 elements per barrier-laden pass and took
// 80 us per array for a 49 k-input window.)
__global__ void __launch_bounds__(1024) k_exclusive_scan2(const uint32_t* __restrict__ in0, uint32_t* __restrict__ out0, const uint32_t* __restrict__ in1,
                                                          uint32_t* __restrict__ out1, size_t n, uint32_t* __restrict__ totals) {
  const uint32_t* in = blockIdx.x ? in1 : in0;
  uint32_t* out = blockIdx.x ? out1 : out0;
  __shared__ uint32_t part[1024];
  const size_t per = (n + 1023) / 1024;
  const size_t a = (size_t)threadIdx.x * per;
  const size_t b = a + per < n ? a + per : n;
  uint32_t s = 0;
  for (size_t i = a; i < b; i++) s += in[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    uint32_t t = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - s;  // exclusive prefix of this thread's chunk
  for (size_t i = a; i < b; i++) {
    uint32_t v = in[i];
    out[i] = run;
    run += v;
  }
  if (threadIdx.x == 1023) totals[blockIdx.x] = part[1023];
}

struct ItemRef { uint32_t input; uint32_t k; };  // which input / which candidate pair

__global__ void k_emit_items(BatchView b, size_t n_inputs, InputPlan* __restrict__ plans, const uint32_t* __restrict__ off_s, const uint32_t* __restrict__ off_e,
                             uint8_t* __restrict__ pk_s, uint8_t* __restrict__ sig_s, ItemRef* __restrict__ ref_s,
                             uint8_t* __restrict__ pk_e, uint8_t* __restrict__ sig_e, ItemRef* __restrict__ ref_e) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inputs) return;
  InputPlan pl = plans[i];
  if (pl.n_items == 0) return;
  bool ecdsa = pl.cls == CLS_P2PK_ECDSA || pl.cls == CLS_MULTISIG_ECDSA;
  uint32_t base = ecdsa ? off_e[i] : off_s[i];
  plans[i].item_base = base;
  const kgv_input& in = b.inputs[i];
  const uint8_t* ss = b.bytes + in.sigscript_off;
  for (uint32_t k = 0; k < pl.n_items; k++) {
    const uint8_t *sig, *key;
    item_location(pl, k, ss, b.entries[i], sig, key);
    size_t it = base + k;
    if (ecdsa) {
      for (int x = 0; x < 33; x++) pk_e[33 * it + x] = key[x];
      for (int x = 0; x < 64; x++) sig_e[64 * it + x] = sig[x];
      ref_e[it] = ItemRef{(uint32_t)i, k};
    } else {
      for (int x = 0; x < 32; x++) pk_s[32 * it + x] = key[x];
      for (int x = 0; x < 64; x++) sig_s[64 * it + x] = sig[x];
      ref_s[it] = ItemRef{(uint32_t)i, k};
    }
  }
}

__global__ void __launch_bounds__(128) k_sighash_reused_v(BatchView b, uint32_t n_txs, const kgv_tx_result* __restrict__ res, SigHashReused* __restrict__ reused) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_txs) return;
  if (res[i].status != KGV_TX_OK) return;
  SigHashReused r;
  sighash_reused(r, b, i);
  reused[i] = r;
}

__global__ void __launch_bounds__(128)
k_item_msgs(BatchView b, const SigHashReused* __restrict__ reused, const uint32_t* __restrict__ input_tx, const InputPlan* __restrict__ plans,
            const ItemRef* __restrict__ refs, size_t n_items, bool ecdsa, uint32_t* __restrict__ msgs) {
  size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= n_items) return;
  ItemRef rf = refs[it];
  InputPlan pl = plans[rf.input];
  const kgv_input& in = b.inputs[rf.input];
  const uint8_t* ss = b.bytes + in.sigscript_off;
  const uint8_t *sig, *key;
  item_location(pl, rf.k, ss, b.entries[rf.input], sig, key);
  uint32_t ht = sig[64];
  uint32_t w[8];
  if (!sighash_type_allowed(ht)) {
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = 0;  // never consulted: resolve reports InvalidSigHashType first
  } else {
    uint32_t tx = input_tx[rf.input];
    SigHashReused r = reused[tx];
    sighash_final(w, b, tx, rf.input, ht, ecdsa, r);
  }
#pragma unroll
  for (int k = 0; k < 8; k++) msgs[8 * it + k] = bswap32(w[k]);
}

__global__ void k_resolve(BatchView b, size_t n_inputs, const uint32_t* __restrict__ input_tx, const kgv_tx_result* __restrict__ res,
                          const InputPlan* __restrict__ plans, const uint8_t* __restrict__ st_s, const uint8_t* __restrict__ st_e,
                          uint8_t* __restrict__ input_err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inputs) return;
  if (res[input_tx[i]].status != KGV_TX_OK) { input_err[i] = KGV_SCRIPT_OK; return; }
  InputPlan pl = plans[i];
  const kgv_input& in = b.inputs[i];
  bool ecdsa = pl.cls == CLS_P2PK_ECDSA || pl.cls == CLS_MULTISIG_ECDSA;
  const uint8_t* st = (ecdsa ? st_e : st_s) + pl.item_base;
  input_err[i] = (uint8_t)resolve_input(pl, b.bytes + in.sigscript_off, b.entries[i], in.sig_op_count, st);
}

__global__ void k_tx_finalize(BatchView b, uint32_t n_txs, const uint8_t* __restrict__ input_err, kgv_tx_result* __restrict__ res) {
  uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
  if (ti >= n_txs) return;
  kgv_tx_result r = res[ti];
  if (r.status != KGV_TX_OK) return;
  const kgv_tx& t = b.txs[ti];
  for (uint32_t i = 0; i < t.n_inputs; i++) {  // check_scripts_sequential order (:170-178)
    uint32_t err = input_err[t.first_input + i];
    if (err == KGV_SCRIPT_OK) continue;
    r.fail_input = i;
    r.script_err = (uint8_t)err;
    if (err == KGV_SCRIPT_NONSTANDARD) r.status = KGV_TX_NEEDS_HOST_VM;
    else r.status = b.inputs[t.first_input + i].sigscript_len == 0 ? KGV_TX_SIGNATURE_EMPTY : KGV_TX_SIGNATURE_INVALID;  // map_script_err :198-200
    break;
  }
  res[ti] = r;
}

__global__ void k_input_tx_index(const kgv_tx* __restrict__ txs, uint32_t n_txs, uint32_t* __restrict__ input_tx) {
  uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
  if (ti >= n_txs) return;
  kgv_tx t = txs[ti];
  for (uint32_t i = 0; i < t.n_inputs; i++) input_tx[t.first_input + i] = ti;
}

// apply accepted transactions to the table (UtxoDiff::add_transaction, utxo_diff.rs:233-247)
__global__ void k_apply_erase(TableView t, const kgv_tx* __restrict__ txs, const kgv_input* __restrict__ inputs, size_t n_inputs,
                              const uint32_t* __restrict__ input_tx, const uint8_t* __restrict__ accept) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inputs) return;
  if (!accept[input_tx[i]]) return;
  uint32_t k[9];
  input_key(k, inputs[i]);
  table_erase(t, k);
}
__global__ void k_output_tx_index(const kgv_tx* __restrict__ txs, uint32_t n_txs, uint32_t* __restrict__ output_tx) {
  uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
  if (ti >= n_txs) return;
  kgv_tx t = txs[ti];
  for (uint32_t i = 0; i < t.n_outputs; i++) output_tx[t.first_output + i] = ti;
}
__global__ void k_apply_insert(TableView t, BatchView b, size_t n_outputs, const uint32_t* __restrict__ output_tx, const uint8_t* __restrict__ accept,
                               const uint64_t* __restrict__ txids, uint64_t pov) {
  size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_outputs) return;
  uint32_t ti = output_tx[o];
  if (!accept[ti]) return;
  const kgv_tx& tx = b.txs[ti];
  const kgv_output& out = b.outputs[o];
  uint32_t k[9];
#pragma unroll
  for (int w = 0; w < 4; w++) { k[2 * w] = (uint32_t)txids[4 * (size_t)ti + w]; k[2 * w + 1] = (uint32_t)(txids[4 * (size_t)ti + w] >> 32); }
  k[8] = (uint32_t)(o - tx.first_output);
  table_put(t, k, out.value, pov, out.spk_version, tx_is_coinbase(tx) ? 1u : 0u, b.bytes + out.script_off, out.script_len);
}
__global__ void __launch_bounds__(128) k_tx_ids_dev(BatchView b, uint32_t n_txs, uint64_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_txs) return;
  uint64_t d[4];
  tx_id(d, b, i);
#pragma unroll
  for (int k = 0; k < 4; k++) out[4 * (size_t)i + k] = d[k];
}

// ---------------------------------------------------------------------------------------------
// K8 MuHash elements (consensus/core/src/muhash.rs:16-33): one element per thread, written straight into level 0
// of its product tree.  Inputs of accepted txs -> denominator tree (the populated entry they spend), outputs ->
// numerator tree (entry = (value, spk, pov_daa_score, tx.is_coinbase)); everything else gets the identity.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_muhash_tx_elements(BatchView b, size_t n_inputs, size_t n_outputs, const uint32_t* __restrict__ input_tx,
                                                            const uint32_t* __restrict__ output_tx, const uint8_t* __restrict__ accept,
                                                            const uint64_t* __restrict__ txids, uint64_t pov, uint32_t* __restrict__ e_den,
                                                            uint32_t* __restrict__ e_num) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_inputs) {
    const uint32_t ti = input_tx[g];
    const DevEntry& e = b.entries[g];
    if (!accept[ti] || !e.found) { u3072_store_one(e_den, n_inputs, g); return; }
    const kgv_input& in = b.inputs[g];
    uint32_t k[8];
#pragma unroll
    for (int w = 0; w < 8; w++)
      k[w] = (uint32_t)in.prev_txid[4 * w] | ((uint32_t)in.prev_txid[4 * w + 1] << 8) | ((uint32_t)in.prev_txid[4 * w + 2] << 16) | ((uint32_t)in.prev_txid[4 * w + 3] << 24);
    uint64_t d[4];
    muhash_utxo_digest(d, k, in.prev_index, e.block_daa_score, e.amount, e.is_coinbase != 0, e.spk_version, e.script, e.script_len);
    muhash_expand_store(e_den, n_inputs, g, d);
    return;
  }
  g -= n_inputs;
  if (g >= n_outputs) return;
  const uint32_t ti = output_tx[g];
  if (!accept[ti]) { u3072_store_one(e_num, n_outputs, g); return; }
  const kgv_tx& tx = b.txs[ti];
  const kgv_output& out = b.outputs[g];
  uint32_t k[8];
#pragma unroll
  for (int w = 0; w < 4; w++) { k[2 * w] = (uint32_t)txids[4 * (size_t)ti + w]; k[2 * w + 1] = (uint32_t)(txids[4 * (size_t)ti + w] >> 32); }
  uint64_t d[4];
  muhash_utxo_digest(d, k, (uint32_t)(g - tx.first_output), pov, out.value, tx_is_coinbase(tx), out.spk_version, b.bytes + out.script_off, out.script_len);
  muhash_expand_store(e_num, n_outputs, g, d);
}
// live entries of table slots [first, first + n) -> elements (empty slots: identity)
__global__ void __launch_bounds__(128) k_muhash_table_elements(TableView t, uint64_t first, size_t n, uint32_t* __restrict__ e_num) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const UtxoSlot* s = &t.slots[first + g];
  if (s->state != SLOT_FULL) { u3072_store_one(e_num, n, g); return; }
  DevEntry e;
  slot_to_entry(e, t, s);
  uint32_t k[8];
#pragma unroll
  for (int w = 0; w < 8; w++) k[w] = s->key[w];
  uint64_t d[4];
  muhash_utxo_digest(d, k, s->key[8], e.block_daa_score, e.amount, e.is_coinbase != 0, e.spk_version, e.script, e.script_len);
  muhash_expand_store(e_num, n, g, d);
}
// contiguous 384-byte values -> level-0 layout
__global__ void k_u3072_scatter(const uint32_t* __restrict__ values, size_t n, uint32_t* __restrict__ e) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  for (int i = 0; i < KGV_U3072_BLOCKS; i++) u3072_store_block(e, n, g, i, values + 96 * g + 8 * i);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline unsigned nblk(size_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

extern "C" int kgv_utxo_create(kgv_ctx* ctx, uint64_t capacity_slots, kgv_utxo_table** out) {
  if (!ctx || !out) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  *out = nullptr;
  CK(cudaSetDevice(ctx->device));
  uint64_t cap = 1024;
  while (cap < capacity_slots) cap <<= 1;
  kgv_utxo_table* t = new kgv_utxo_table();
  t->mask = cap - 1;
  t->overflow_cap = cap * 8 < (64ull << 20) ? (64ull << 20) : cap * 8;
  cudaError_t e = cudaMalloc((void**)&t->slots, cap * sizeof(UtxoSlot));
  if (e == cudaSuccess) e = cudaMalloc((void**)&t->overflow, t->overflow_cap);
  if (e == cudaSuccess) e = cudaMalloc((void**)&t->counters, 16 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    ctx->err = std::string("cudaMalloc failed for the UTXO table: ") + cudaGetErrorString(e);
    (void)cudaGetLastError();
    if (t->slots) cudaFree(t->slots);
    if (t->overflow) cudaFree(t->overflow);
    delete t;
    return KGV_ERR_NOMEM;
  }
  CK(cudaMemsetAsync(t->slots, 0, cap * sizeof(UtxoSlot), ctx->stream));
  CK(cudaMemsetAsync(t->counters, 0, 16 * sizeof(unsigned long long), ctx->stream));
  CK(cudaMalloc((void**)&t->d_view, sizeof(TableView)));
  TableView hv = view_of(t);
  CK(cudaMemcpyAsync(t->d_view, &hv, sizeof hv, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *out = t;
  return KGV_OK;
}

// ---------------------------------------------------------------------------------------------
// composed views: a diff layer on the device (utxo_view.rs:22-35, utxo_diff.rs:15-19)
// ---------------------------------------------------------------------------------------------
extern "C" int kgv_utxo_view_create(kgv_ctx* ctx, kgv_utxo_table* base, uint64_t capacity_slots, kgv_utxo_table** out) {
  if (!ctx || !base || !out) return KGV_ERR_ARG;
  int rc = kgv_utxo_create(ctx, capacity_slots, out);
  if (rc) return rc;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  (*out)->base = base;
  TableView hv = view_of(*out);  // now with `below`
  CK(cudaMemcpyAsync((*out)->d_view, &hv, sizeof hv, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return KGV_OK;
}
// write_diff_batch (utxo_set.rs:107-112): delete what the layer removed, put what it added - applied to the layer below through that layer's own
// view semantics (so a stack of diffs folds downwards one level at a time)
__global__ void k_view_commit_removes(TableView top, TableView below) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > top.mask) return;
  const UtxoSlot* s = &top.slots[i];
  const uint32_t st = s->state;
  if (st != SLOT_REMOVED && st != SLOT_FULLH) return;
  uint32_t k[9];
#pragma unroll
  for (int w = 0; w < 9; w++) k[w] = s->key[w];
  table_erase(below, k);
}
__global__ void k_view_commit_adds(TableView top, TableView below) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > top.mask) return;
  const UtxoSlot* s = &top.slots[i];
  const uint32_t st = s->state;
  if (st != SLOT_FULL && st != SLOT_FULLH) return;
  SlotHead h;
  slot_load_head(h, s);
  DevEntry e;
  TableView solo = top;
  solo.below = nullptr;
  head_to_entry(e, solo, s, h);
  uint32_t k[9];
#pragma unroll
  for (int w = 0; w < 9; w++) k[w] = s->key[w];
  table_put(below, k, e.amount, e.block_daa_score, e.spk_version, e.is_coinbase, e.script, e.script_len);
}
static int view_clear(kgv_ctx* ctx, kgv_utxo_table* v) {
  CK(cudaMemsetAsync(v->slots, 0, (v->mask + 1) * sizeof(UtxoSlot), ctx->stream));
  CK(cudaMemsetAsync(v->counters, 0, 16 * sizeof(unsigned long long), ctx->stream));
  return KGV_OK;
}
extern "C" int kgv_utxo_view_commit(kgv_ctx* ctx, kgv_utxo_table* view) {
  if (!ctx || !view || !view->base) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  TableView top = view_of(view), below = view_of(view->base);
  k_view_commit_removes<<<nblk(view->mask + 1, 128), 128, 0, ctx->stream>>>(top, below);
  CK(cudaGetLastError());
  k_view_commit_adds<<<nblk(view->mask + 1, 128), 128, 0, ctx->stream>>>(top, below);
  CK(cudaGetLastError());
  ctx->launches += 2;
  return view_clear(ctx, view);
}
extern "C" int kgv_utxo_view_discard(kgv_ctx* ctx, kgv_utxo_table* view) {
  if (!ctx || !view || !view->base) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  return view_clear(ctx, view);
}
extern "C" void kgv_utxo_destroy(kgv_ctx* ctx, kgv_utxo_table* t) {
  if (!t) return;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  cudaFree(t->slots); cudaFree(t->overflow); cudaFree(t->counters);
  if (t->d_view) cudaFree(t->d_view);
  delete t;
}

// stage a host array on the device inside ctx->d_in at a running offset
struct Stager {
  kgv_ctx* ctx;
  size_t off = 0;
  explicit Stager(kgv_ctx* c) : ctx(c) {}
};

extern "C" int kgv_utxo_lookup(kgv_ctx* ctx, kgv_utxo_table* t, const uint8_t* keys36, size_t n, kgv_utxo_entry* entries, uint8_t* scripts_out,
                               uint32_t script_stride, uint8_t* found) {
  if (!ctx || !t) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (n == 0) return KGV_OK;
  if (!keys36 || !entries || !found || (script_stride && !scripts_out)) { ctx->err = "null buffer"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  bool dev = kgv_ptr_is_device(keys36);
  const uint8_t* dk = keys36;
  kgv_utxo_entry* de = entries;
  uint8_t *ds = scripts_out, *df = found;
  size_t o_e = 0, o_s = al256(n * sizeof(kgv_utxo_entry)), o_f = al256(o_s + n * (size_t)script_stride);
  if (!dev) {
    int rc = kgv_reserve(ctx, &ctx->d_in, &ctx->d_in_cap, n * 36);
    if (rc) return rc;
    rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, o_f + n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_in, keys36, n * 36, cudaMemcpyHostToDevice, ctx->stream));
    dk = ctx->d_in; de = (kgv_utxo_entry*)(ctx->d_out + o_e); ds = ctx->d_out + o_s; df = ctx->d_out + o_f;
  }
  k_utxo_lookup<<<nblk(n, 128), 128, 0, ctx->stream>>>(view_of(t), dk, n, de, ds, script_stride, df);
  CK(cudaGetLastError());
  ctx->launches++;
  if (!dev) {
    CK(cudaMemcpyAsync(entries, de, n * sizeof(kgv_utxo_entry), cudaMemcpyDeviceToHost, ctx->stream));
    if (script_stride) CK(cudaMemcpyAsync(scripts_out, ds, n * (size_t)script_stride, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(found, df, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return KGV_OK;
}

extern "C" int kgv_utxo_apply_diff(kgv_ctx* ctx, kgv_utxo_table* t, const uint8_t* rem_keys36, size_t n_rem, uint8_t* rem_status,
                                   const uint8_t* add_keys36, const kgv_utxo_entry* add_entries, const uint8_t* add_bytes, size_t n_add_bytes, size_t n_add,
                                   uint8_t* add_status) {
  if (!ctx || !t) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if ((n_rem && !rem_keys36) || (n_add && (!add_keys36 || !add_entries || (n_add_bytes && !add_bytes)))) { ctx->err = "null buffer"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const void* probe = n_rem ? (const void*)rem_keys36 : (const void*)add_keys36;
  if (!probe) return KGV_OK;
  bool dev = kgv_ptr_is_device(probe);
  size_t o_rk = 0, o_ak = al256(n_rem * 36), o_ae = al256(o_ak + n_add * 36), o_ab = al256(o_ae + n_add * sizeof(kgv_utxo_entry));
  size_t o_rs = 0, o_as = al256(n_rem);
  const uint8_t *drk = rem_keys36, *dak = add_keys36, *dab = add_bytes;
  const kgv_utxo_entry* dae = add_entries;
  uint8_t *drs = rem_status, *das = add_status;
  if (!dev) {
    int rc = kgv_reserve(ctx, &ctx->d_in, &ctx->d_in_cap, o_ab + n_add_bytes + 16);
    if (rc) return rc;
    rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, o_as + n_add + 16);
    if (rc) return rc;
    if (n_rem) CK(cudaMemcpyAsync(ctx->d_in + o_rk, rem_keys36, n_rem * 36, cudaMemcpyHostToDevice, ctx->stream));
    if (n_add) {
      CK(cudaMemcpyAsync(ctx->d_in + o_ak, add_keys36, n_add * 36, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(ctx->d_in + o_ae, add_entries, n_add * sizeof(kgv_utxo_entry), cudaMemcpyHostToDevice, ctx->stream));
      if (n_add_bytes) CK(cudaMemcpyAsync(ctx->d_in + o_ab, add_bytes, n_add_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    drk = ctx->d_in + o_rk; dak = ctx->d_in + o_ak; dae = (const kgv_utxo_entry*)(ctx->d_in + o_ae); dab = ctx->d_in + o_ab;
    drs = ctx->d_out + o_rs; das = ctx->d_out + o_as;
  }
  if (n_rem) { k_utxo_erase<<<nblk(n_rem, 128), 128, 0, ctx->stream>>>(view_of(t), drk, n_rem, drs); CK(cudaGetLastError()); ctx->launches++; }
  if (n_add) { k_utxo_insert<<<nblk(n_add, 128), 128, 0, ctx->stream>>>(view_of(t), dak, dae, dab, n_add, das); CK(cudaGetLastError()); ctx->launches++; }
  if (!dev) {
    if (n_rem && rem_status) CK(cudaMemcpyAsync(rem_status, drs, n_rem, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_add && add_status) CK(cudaMemcpyAsync(add_status, das, n_add, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return KGV_OK;
}

extern "C" int kgv_utxo_export(kgv_ctx* ctx, kgv_utxo_table* t, uint8_t* keys36, kgv_utxo_entry* entries, uint8_t* bytes, size_t max_n, size_t bytes_cap, size_t* n_out,
                               size_t* bytes_out) {
  if (!ctx || !t) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (t->base) { ctx->err = "export is defined on plain tables: commit the view first"; return KGV_ERR_ARG; }
  const bool counting = !keys36;
  if (!counting && (!entries || (bytes_cap && !bytes))) { ctx->err = "null buffer"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const bool dev = counting ? true : kgv_ptr_is_device(keys36);
  uint8_t *dk = keys36, *db = bytes;
  kgv_utxo_entry* de = entries;
  size_t o_e = al256(max_n * 36), o_b = al256(o_e + max_n * sizeof(kgv_utxo_entry));
  if (!counting && !dev) {
    int rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, o_b + bytes_cap + 16);
    if (rc) return rc;
    dk = ctx->d_out; de = (kgv_utxo_entry*)(ctx->d_out + o_e); db = ctx->d_out + o_b;
  }
  unsigned long long* cnt = t->counters + 8;  // (the digest's scratch words)
  CK(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), ctx->stream));
  k_utxo_export<<<nblk(t->mask + 1, 128), 128, 0, ctx->stream>>>(view_of(t), counting ? nullptr : dk, de, db, max_n, bytes_cap, cnt);
  CK(cudaGetLastError());
  ctx->launches++;
  unsigned long long c[2];
  CK(cudaMemcpyAsync(c, cnt, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (n_out) *n_out = (size_t)c[0];
  if (bytes_out) *bytes_out = (size_t)c[1];
  if (counting) return KGV_OK;
  if (c[0] > max_n || c[1] > bytes_cap) { ctx->err = "kgv_utxo_export: the caller's arrays are too small (sizes returned)"; return KGV_ERR_NOMEM; }
  if (!dev) {
    if (c[0]) {
      CK(cudaMemcpyAsync(keys36, dk, (size_t)c[0] * 36, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaMemcpyAsync(entries, de, (size_t)c[0] * sizeof(kgv_utxo_entry), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (c[1]) CK(cudaMemcpyAsync(bytes, db, (size_t)c[1], cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return KGV_OK;
}

extern "C" int kgv_utxo_import_chunk(kgv_ctx* ctx, kgv_utxo_table* t, const uint8_t* keys36, const kgv_utxo_entry* entries, const uint8_t* bytes, size_t n_bytes, size_t n,
                                     uint8_t* numerator384) {
  if (!ctx || !t) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!numerator384) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  if (n == 0) return KGV_OK;
  if (!keys36 || !entries || (n_bytes && !bytes)) { ctx->err = "null buffer"; return KGV_ERR_ARG; }
  if (kgv_ptr_is_device(numerator384)) { ctx->err = "numerator384 is a host value"; return KGV_ERR_ARG; }
  for (size_t i = 0; !kgv_ptr_is_device(keys36) && i < n; i++)
    if ((uint64_t)entries[i].script_off + entries[i].script_len > n_bytes) { ctx->err = "entry script outside the byte arena"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const uint8_t *dk = keys36, *db = bytes;
  const kgv_utxo_entry* de = entries;
  if (!kgv_ptr_is_device(keys36)) {  // one upload serves the insert and the multiset
    size_t o_e = al256(n * 36), o_b = al256(o_e + n * sizeof(kgv_utxo_entry));
    int rc = kgv_reserve(ctx, &ctx->d_scratch, &ctx->d_scratch_cap, o_b + n_bytes + 16);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_scratch, keys36, n * 36, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_scratch + o_e, entries, n * sizeof(kgv_utxo_entry), cudaMemcpyHostToDevice, ctx->stream));
    if (n_bytes) CK(cudaMemcpyAsync(ctx->d_scratch + o_b, bytes, n_bytes, cudaMemcpyHostToDevice, ctx->stream));
    dk = ctx->d_scratch; de = (const kgv_utxo_entry*)(ctx->d_scratch + o_e); db = ctx->d_scratch + o_b;
  }
  // pruning_meta.utxo_set.write_many(chunk) (consensus/mod.rs:1072)
  int rc = kgv_utxo_apply_diff(ctx, t, nullptr, 0, nullptr, dk, de, db, n_bytes, n, nullptr);
  if (rc) return rc;
  // chunk.par_iter().map(MuHash::from_utxo).reduce(combine) (:1075-1080), then current_multiset.combine (:1082)
  uint32_t *e_den = nullptr, *e_num = nullptr;
  rc = kgv_mu_reserve(ctx, 0, n, &e_den, &e_num);
  if (rc) return rc;
  k_muhash_utxo_elements<<<nblk(n, 128), 128, 0, ctx->stream>>>(dk, de, db, n, e_num);
  CK(cudaGetLastError());
  ctx->launches++;
  uint8_t chunk_num[384], chunk_den[384], one[384];
  rc = kgv_mu_reduce(ctx, 0, n, chunk_num, chunk_den);
  if (rc) return rc;
  memset(one, 0, sizeof one);
  one[0] = 1;
  uint8_t den[384];
  memcpy(den, one, sizeof den);
  return kgv_muhash_combine(ctx, numerator384, den, chunk_num, one);
}

extern "C" int kgv_utxo_count(kgv_ctx* ctx, kgv_utxo_table* t, uint64_t* count) {
  if (!ctx || !t || !count) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (t->base) { ctx->err = "count / digest / MuHash are defined on plain tables: commit the view first"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  unsigned long long c[4];
  CK(cudaMemcpyAsync(c, t->counters, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (c[3]) { ctx->err = "UTXO table insert failures (table or overflow arena full)"; return KGV_ERR_NOMEM; }
  *count = c[0];
  return KGV_OK;
}

extern "C" int kgv_utxo_digest(kgv_ctx* ctx, kgv_utxo_table* t, uint8_t out32[32]) {
  if (!ctx || !t || !out32) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (t->base) { ctx->err = "count / digest / MuHash are defined on plain tables: commit the view first"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  unsigned long long* acc = t->counters + 8;
  CK(cudaMemsetAsync(acc, 0, 8 * sizeof(unsigned long long), ctx->stream));
  k_utxo_digest<<<nblk(t->mask + 1, 128), 128, 0, ctx->stream>>>(view_of(t), acc);
  CK(cudaGetLastError());
  ctx->launches++;
  unsigned long long h[8];
  CK(cudaMemcpyAsync(h, acc, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  unsigned __int128 c = 0;
  for (int i = 0; i < 8; i++) {  // fold the 32-bit limb sums into a 256-bit little-endian integer (mod 2^256)
    c += h[i];
    uint32_t limb = (uint32_t)c;
    c >>= 32;
    for (int b = 0; b < 4; b++) out32[4 * i + b] = (uint8_t)(limb >> (8 * b));
  }
  return KGV_OK;
}

// The script phase of check_scripts (tx_validation_in_utxo_context.rs:162-200) for every transaction whose dres[].status is
// KGV_TX_OK: plan -> scan -> emit -> sighash -> verify -> resolve -> finalize, all enqueued on ctx->stream (the ECDSA items on
// the side stream).  v.entries must be populated.  Uses ctx->d_scratch (plans, counts, offsets, sub-hashes) and ctx->d_in
// (item arrays); one host synchronisation (the two item totals).
int kgv_scripts_phase(kgv_ctx* ctx, const BatchView& v, size_t nt, size_t ni, const uint32_t* itx, kgv_tx_result* dres, uint64_t* n_items_out) {
  if (n_items_out) *n_items_out = 0;
  if (ni == 0) return KGV_OK;
  size_t o_plan = 0;
  size_t o_cs = al256(o_plan + ni * sizeof(InputPlan));
  size_t o_ce = al256(o_cs + ni * 4);
  size_t o_os = al256(o_ce + ni * 4);
  size_t o_oe = al256(o_os + ni * 4);
  size_t o_tot = al256(o_oe + ni * 4);
  size_t o_err = al256(o_tot + 64);
  size_t o_reu = al256(o_err + ni);
  size_t total = al256(o_reu + nt * sizeof(SigHashReused));
  int rc = kgv_reserve(ctx, &ctx->d_scratch, &ctx->d_scratch_cap, total);
  if (rc) return rc;
  uint8_t* S = ctx->d_scratch;
  InputPlan* plans = (InputPlan*)(S + o_plan);
  uint32_t *cs = (uint32_t*)(S + o_cs), *ce = (uint32_t*)(S + o_ce), *os = (uint32_t*)(S + o_os), *oe = (uint32_t*)(S + o_oe), *tot = (uint32_t*)(S + o_tot);
  uint8_t* ierr = S + o_err;
  SigHashReused* reu = (SigHashReused*)(S + o_reu);
  cudaStream_t st = ctx->stream;
  k_plan<<<nblk(ni, 128), 128, 0, st>>>(v, ni, itx, dres, plans, cs, ce);
  CK(cudaGetLastError());
  STAGE("plan");
  k_exclusive_scan2<<<2, 1024, 0, st>>>(cs, os, ce, oe, ni, tot);
  CK(cudaGetLastError());
  ctx->launches += 2;
  uint32_t totals[2];
  CK(cudaMemcpyAsync(totals, tot, sizeof totals, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  size_t ns = totals[0], ne = totals[1];
  if (n_items_out) *n_items_out = ns + ne;
  if (kgv_debug_on()) fprintf(stderr, "[kgv] items: schnorr %zu ecdsa %zu\n", ns, ne);
  // multi-GPU (kgv_set_sharding): the candidate pairs are split into n_ranks contiguous ranges; this rank verifies one and the
  // status bytes are exchanged before the scripts are resolved.  per_* = range length (the status arrays are padded to nr * per)
  int nr = 1, rk = 0;
  if (ctx->shard_comm) nr = kgv_comm_ranks(ctx->shard_comm, &rk);
  auto per_of = [&](size_t n) { size_t p = (n + nr - 1) / nr; return (p + 255) & ~(size_t)255; };
  const size_t per_s = nr > 1 ? per_of(ns) : ns, per_e = nr > 1 ? per_of(ne) : ne;
  const size_t s_lo = nr > 1 ? (rk * per_s < ns ? rk * per_s : ns) : 0, s_hi = nr > 1 ? ((rk + 1) * per_s < ns ? (rk + 1) * per_s : ns) : ns;
  const size_t e_lo = nr > 1 ? (rk * per_e < ne ? rk * per_e : ne) : 0, e_hi = nr > 1 ? ((rk + 1) * per_e < ne ? (rk + 1) * per_e : ne) : ne;
  // item arrays live in d_in (pk/sig/msg, status, refs)
  size_t i_pks = 0, i_sigs = al256(i_pks + ns * 32), i_msgs = al256(i_sigs + ns * 64), i_refs = al256(i_msgs + ns * 32), i_sts = al256(i_refs + ns * sizeof(ItemRef));
  size_t i_pke = al256(i_sts + (nr > 1 ? nr * per_s : ns)), i_sige = al256(i_pke + ne * 33), i_msge = al256(i_sige + ne * 64), i_refe = al256(i_msge + ne * 32),
         i_ste = al256(i_refe + ne * sizeof(ItemRef));
  // signature cache (kgv_set_sigcache; not combined with sharding): digests, miss lists and miss counts of the two item kinds
  kgv_sigcache* sc = nr == 1 ? ctx->sigcache : nullptr;
  size_t i_digs = al256(i_ste + (nr > 1 ? nr * per_e : ne) + 64), i_idxs = al256(i_digs + (sc ? ns * 32 : 0)), i_dige = al256(i_idxs + (sc ? ns * 4 : 0)),
         i_idxe = al256(i_dige + (sc ? ne * 32 : 0)), i_nm = al256(i_idxe + (sc ? ne * 4 : 0));
  size_t total2 = al256(i_nm + 64);
  rc = kgv_reserve(ctx, &ctx->d_in, &ctx->d_in_cap, total2);
  if (rc) return rc;
  uint8_t* I = ctx->d_in;
  if (ns + ne) {
    k_emit_items<<<nblk(ni, 128), 128, 0, st>>>(v, ni, plans, os, oe, I + i_pks, I + i_sigs, (ItemRef*)(I + i_refs), I + i_pke, I + i_sige, (ItemRef*)(I + i_refe));
    CK(cudaGetLastError());
    k_sighash_reused_v<<<nblk(nt, 128), 128, 0, st>>>(v, (uint32_t)nt, dres, reu);
    CK(cudaGetLastError());
    ctx->launches += 2;
    STAGE("emit+reused");
  }
  if (ns && ne) CK(cudaEventRecord(ctx->ev_fork, st));  // fork point: everything both item kinds depend on is queued
  if (ns) {
    if (s_hi > s_lo) {
      k_item_msgs<<<nblk(s_hi - s_lo, 128), 128, 0, st>>>(v, reu, itx, plans, (const ItemRef*)(I + i_refs) + s_lo, s_hi - s_lo, false, (uint32_t*)(I + i_msgs) + 8 * s_lo);
      CK(cudaGetLastError());
      ctx->launches++;
      STAGE("msgs schnorr");
      if (sc) {  // hits are answered from the table, only the misses are verified (and remembered)
        rc = kgv_sigcache_lookup(ctx, sc, I + i_pks, I + i_msgs, I + i_sigs, ns, false, I + i_sts, I + i_digs, (uint32_t*)(I + i_idxs), (uint32_t*)(I + i_nm), st);
        if (rc) return rc;
        rc = kgv_launch_verify(ctx, I + i_pks, I + i_msgs, I + i_sigs, ns, I + i_sts, false, nullptr, false, (const uint32_t*)(I + i_idxs), (const uint32_t*)(I + i_nm));
        if (rc) return rc;
        rc = kgv_sigcache_insert(ctx, sc, I + i_sts, I + i_digs, (const uint32_t*)(I + i_idxs), (const uint32_t*)(I + i_nm), ns, st);
        if (rc) return rc;
      } else {
        rc = kgv_launch_verify(ctx, I + i_pks + 32 * s_lo, I + i_msgs + 32 * s_lo, I + i_sigs + 64 * s_lo, s_hi - s_lo, I + i_sts + s_lo, false);
        if (rc) return rc;
      }
      STAGE("verify schnorr");
    }
  }
  if (ne) {
    // with both kinds present the ECDSA items run on the side stream so the two (often sub-wave) verify
    // launches share the SMs instead of queueing behind each other
    const bool fork = ns != 0 && !kgv_debug_on();
    cudaStream_t se = fork ? ctx->aux_stream : st;
    if (fork) CK(cudaStreamWaitEvent(se, ctx->ev_fork, 0));
    if (e_hi > e_lo) {
      k_item_msgs<<<nblk(e_hi - e_lo, 128), 128, 0, se>>>(v, reu, itx, plans, (const ItemRef*)(I + i_refe) + e_lo, e_hi - e_lo, true, (uint32_t*)(I + i_msge) + 8 * e_lo);
      CK(cudaGetLastError());
      ctx->launches++;
      STAGE("msgs ecdsa");
      if (sc) {
        rc = kgv_sigcache_lookup(ctx, sc, I + i_pke, I + i_msge, I + i_sige, ne, true, I + i_ste, I + i_dige, (uint32_t*)(I + i_idxe), (uint32_t*)(I + i_nm) + 1, se);
        if (rc) return rc;
        rc = kgv_launch_verify(ctx, I + i_pke, I + i_msge, I + i_sige, ne, I + i_ste, true, se, true, (const uint32_t*)(I + i_idxe), (const uint32_t*)(I + i_nm) + 1);
        if (rc) return rc;
        rc = kgv_sigcache_insert(ctx, sc, I + i_ste, I + i_dige, (const uint32_t*)(I + i_idxe), (const uint32_t*)(I + i_nm) + 1, ne, se);
        if (rc) return rc;
      } else {
        rc = kgv_launch_verify(ctx, I + i_pke + 33 * e_lo, I + i_msge + 32 * e_lo, I + i_sige + 64 * e_lo, e_hi - e_lo, I + i_ste + e_lo, true, se, true);
        if (rc) return rc;
      }
    }
    if (fork) {
      CK(cudaEventRecord(ctx->ev_join, se));
      CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    }
    STAGE("verify ecdsa");
  }
  if (nr > 1) {
    if (ns) { rc = kgv_comm_exchange_slices(ctx, ctx->shard_comm, I + i_sts, per_s); if (rc) return rc; }
    if (ne) { rc = kgv_comm_exchange_slices(ctx, ctx->shard_comm, I + i_ste, per_e); if (rc) return rc; }
    STAGE("verdict exchange");
  }
  k_resolve<<<nblk(ni, 128), 128, 0, st>>>(v, ni, itx, dres, plans, I + i_sts, I + i_ste, ierr);
  CK(cudaGetLastError());
  STAGE("resolve");
  k_tx_finalize<<<nblk(nt, 128), 128, 0, st>>>(v, (uint32_t)nt, ierr, dres);
  CK(cudaGetLastError());
  ctx->launches += 2;
  return KGV_OK;
}

// ---------------------------------------------------------------------------------------------
// Non-standard scripts behind the table: the host script engine needs the populated entries, which only the library can
// see.  For every transaction reported KGV_TX_NEEDS_HOST_VM the entries the device populated (dent) are gathered at their
// TRUE script length (no stride, no truncation), a host-resident populated batch is assembled and kgv_check_scripts_host
// decides (its signature checks go back to the GPU in batches).  `dres` is patched in place (device array of n_txs).
// Rare path: a few synchronous copies.
// ---------------------------------------------------------------------------------------------
__global__ void k_gather_entry_meta(const DevEntry* __restrict__ dent, const uint32_t* __restrict__ list, uint32_t n, kgv_utxo_entry* __restrict__ out) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  DevEntry d = dent[list[j]];
  kgv_utxo_entry e;
  memset(&e, 0, sizeof e);
  e.amount = d.amount; e.block_daa_score = d.block_daa_score; e.script_len = d.found ? d.script_len : 0; e.spk_version = d.spk_version; e.is_coinbase = d.is_coinbase;
  e.pad_[0] = d.found ? 0 : 1;
  out[j] = e;
}
__global__ void k_gather_entry_scripts(const DevEntry* __restrict__ dent, const uint32_t* __restrict__ list, uint32_t n, const uint64_t* __restrict__ off,
                                       uint8_t* __restrict__ out) {
  uint32_t j = blockIdx.x;
  if (j >= n) return;
  DevEntry d = dent[list[j]];
  if (!d.found) return;
  for (uint32_t b = threadIdx.x; b < d.script_len; b += blockDim.x) out[off[j] + b] = d.script[b];
}

int kgv_host_vm_resolve(kgv_ctx* ctx, const kgv_tx_batch* batch, const kgv_dev_batch& d, const DevEntry* dent, kgv_tx_result* dres) {
  const size_t nt = d.n_txs, ni = d.n_inputs;
  cudaStream_t st = ctx->stream;
  std::vector<kgv_tx_result> hres(nt);
  CK(cudaMemcpyAsync(hres.data(), dres, nt * sizeof(kgv_tx_result), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  std::vector<uint32_t> idx;
  for (size_t i = 0; i < nt; i++) if (hres[i].status == KGV_TX_NEEDS_HOST_VM) idx.push_back((uint32_t)i);
  if (idx.empty()) return KGV_OK;
  // host copy of the batch
  std::vector<kgv_tx> htx; std::vector<kgv_input> hin; std::vector<kgv_output> hout; std::vector<uint8_t> hby;
  const kgv_tx* txs = batch->txs; const kgv_input* ins = batch->inputs; const kgv_output* outs = batch->outputs; const uint8_t* by = batch->bytes;
  const void* probe = batch->n_txs ? (const void*)batch->txs : (const void*)batch->bytes;
  if (probe && kgv_ptr_is_device(probe)) {
    htx.resize(nt); hin.resize(ni); hout.resize(d.n_outputs); hby.resize(d.n_bytes);
    CK(cudaMemcpyAsync(htx.data(), d.txs, nt * sizeof(kgv_tx), cudaMemcpyDeviceToHost, st));
    if (ni) CK(cudaMemcpyAsync(hin.data(), d.inputs, ni * sizeof(kgv_input), cudaMemcpyDeviceToHost, st));
    if (d.n_outputs) CK(cudaMemcpyAsync(hout.data(), d.outputs, d.n_outputs * sizeof(kgv_output), cudaMemcpyDeviceToHost, st));
    if (d.n_bytes) CK(cudaMemcpyAsync(hby.data(), d.bytes, d.n_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    txs = htx.data(); ins = hin.data(); outs = hout.data(); by = hby.data();
  }
  std::vector<uint32_t> list;
  for (uint32_t ti : idx)
    for (uint32_t k = 0; k < txs[ti].n_inputs; k++) list.push_back(txs[ti].first_input + k);
  const size_t nl = list.size();
  std::vector<kgv_utxo_entry> ents(ni);
  for (auto& e : ents) { memset(&e, 0, sizeof e); e.pad_[0] = 1; }
  std::vector<uint8_t> arena(by, by + d.n_bytes);
  if (nl) {
    size_t o_list = 0, o_meta = al256(nl * 4), o_off = al256(o_meta + nl * sizeof(kgv_utxo_entry));
    int rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, al256(o_off + nl * 8));
    if (rc) return rc;
    uint8_t* O = ctx->d_out;
    CK(cudaMemcpyAsync(O + o_list, list.data(), nl * 4, cudaMemcpyHostToDevice, st));
    k_gather_entry_meta<<<nblk(nl, 128), 128, 0, st>>>(dent, (const uint32_t*)(O + o_list), (uint32_t)nl, (kgv_utxo_entry*)(O + o_meta));
    CK(cudaGetLastError());
    ctx->launches++;
    std::vector<kgv_utxo_entry> meta(nl);
    CK(cudaMemcpyAsync(meta.data(), O + o_meta, nl * sizeof(kgv_utxo_entry), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<uint64_t> off(nl);
    uint64_t run = 0;
    for (size_t j = 0; j < nl; j++) { off[j] = run; run += (meta[j].script_len + 7u) & ~7u; }
    if (d.n_bytes + run > 0xFFFFFFF0ull) { ctx->err = "populated scripts do not fit a 32-bit arena offset"; return KGV_ERR_ARG; }
    std::vector<uint8_t> scripts(run + 8);
    if (run) {
      // d_scratch is free again at this point (the script phase of this call has completed)
      rc = kgv_reserve(ctx, &ctx->d_scratch, &ctx->d_scratch_cap, run + 256);
      if (rc) return rc;
      CK(cudaMemcpyAsync(O + o_off, off.data(), nl * 8, cudaMemcpyHostToDevice, st));
      k_gather_entry_scripts<<<(unsigned)nl, 64, 0, st>>>(dent, (const uint32_t*)(O + o_list), (uint32_t)nl, (const uint64_t*)(O + o_off), ctx->d_scratch);
      CK(cudaGetLastError());
      ctx->launches++;
      CK(cudaMemcpyAsync(scripts.data(), ctx->d_scratch, run, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
    const size_t base = arena.size();
    arena.insert(arena.end(), scripts.begin(), scripts.end());
    for (size_t j = 0; j < nl; j++) {
      kgv_utxo_entry e = meta[j];
      e.script_off = (uint32_t)(base + off[j]);
      ents[list[j]] = e;
    }
  }
  // a transaction with an absent entry never reaches the script phase (MissingTxOutpoints comes first), so every listed entry is present
  kgv_tx_batch hb;
  hb.txs = txs; hb.n_txs = nt; hb.inputs = ins; hb.n_inputs = ni; hb.outputs = outs; hb.n_outputs = d.n_outputs; hb.entries = ents.data();
  hb.bytes = arena.data(); hb.n_bytes = arena.size();
  std::vector<kgv_tx_result> out(idx.size());
  int rc = kgv_check_scripts_host(ctx, &hb, idx.data(), idx.size(), out.data());
  if (rc) return rc;
  for (size_t j = 0; j < idx.size(); j++) {
    kgv_tx_result& r = hres[idx[j]];
    r.status = out[j].status; r.script_err = out[j].script_err; r.fail_input = out[j].fail_input;
  }
  CK(cudaMemcpyAsync(dres, hres.data(), nt * sizeof(kgv_tx_result), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  return KGV_OK;
}
__global__ void k_count_status(const kgv_tx_result* __restrict__ res, uint32_t n, uint8_t what, unsigned long long* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = i < n && res[i].status == what;
  unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

// shared core of kgv_validate_populated / kgv_validate_txs
static int validate_core(kgv_ctx* ctx, kgv_utxo_table* table, const kgv_tx_batch* batch, uint64_t pov, uint32_t flags, const kgv_params* prm,
                         kgv_tx_result* results) {
  if (!batch || !prm || (batch->n_txs && !results)) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  if (flags > KGV_FLAGS_SCRIPTS_ONLY) { ctx->err = "unknown validation flags"; return KGV_ERR_ARG; }
  if (batch->n_txs == 0) return KGV_OK;
  CK(cudaSetDevice(ctx->device));
  kgv_dev_batch d;
  int rc = kgv_batch_to_device(ctx, batch, &d, table == nullptr);
  if (rc) return rc;
  const size_t nt = d.n_txs, ni = d.n_inputs;
  size_t o_ent = 0;
  size_t o_itx = al256(o_ent + ni * sizeof(DevEntry));
  size_t o_res = al256(o_itx + ni * 4);
  rc = kgv_reserve(ctx, &ctx->d_work, &ctx->d_work_cap, al256(o_res + nt * sizeof(kgv_tx_result)));
  if (rc) return rc;
  uint8_t* S = ctx->d_work;
  DevEntry* dent = (DevEntry*)(S + o_ent);
  uint32_t* itx = (uint32_t*)(S + o_itx);
  kgv_tx_result* dres = (kgv_tx_result*)(S + o_res);
  cudaStream_t st = ctx->stream;

  if (ni) {
    if (table) k_populate<<<nblk(ni, 128), 128, 0, st>>>(view_of(table), d.inputs, ni, dent);
    else k_entries_from_batch<<<nblk(ni, 128), 128, 0, st>>>(d.entries, d.bytes, ni, dent);
    CK(cudaGetLastError());
    k_input_tx_index<<<nblk(nt, 128), 128, 0, st>>>(d.txs, (uint32_t)nt, itx);
    CK(cudaGetLastError());
    ctx->launches += 2;
  }
  STAGE("populate");
  BatchView v{d.txs, d.inputs, d.outputs, dent, d.bytes};
  k_tx_context<<<nblk(nt, 128), 128, 0, st>>>(v, (uint32_t)nt, pov, flags, *prm, dres);
  CK(cudaGetLastError());
  ctx->launches++;
  STAGE("tx_context");
  if (flags != KGV_FLAGS_SKIP_SCRIPT_CHECKS && ni) {
    rc = kgv_scripts_phase(ctx, v, nt, ni, itx, dres, nullptr);
    if (rc) return rc;
    if (table) {
      // utxo_validation.rs:282-309 accepts ANY transaction whose scripts execute successfully: a non-standard spend must not
      // leave this call undecided, and only the library can read the entries it was populated with
      unsigned long long* cnt = table->counters + 4;
      unsigned long long n_vm = 0;
      CK(cudaMemsetAsync(cnt, 0, 8, st));
      k_count_status<<<nblk(nt, 256), 256, 0, st>>>(dres, (uint32_t)nt, KGV_TX_NEEDS_HOST_VM, cnt);
      CK(cudaGetLastError());
      ctx->launches++;
      CK(cudaMemcpyAsync(&n_vm, cnt, 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (n_vm) {
        rc = kgv_host_vm_resolve(ctx, batch, d, dent, dres);
        if (rc) return rc;
      }
    }
  }
  if (kgv_ptr_is_device(results)) {
    CK(cudaMemcpyAsync(results, dres, nt * sizeof(kgv_tx_result), cudaMemcpyDeviceToDevice, st));
  } else {
    CK(cudaMemcpyAsync(results, dres, nt * sizeof(kgv_tx_result), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return KGV_OK;
}

extern "C" int kgv_validate_populated(kgv_ctx* ctx, const kgv_tx_batch* batch, uint64_t pov_daa_score, uint32_t flags, const kgv_params* params,
                                      kgv_tx_result* results) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  return validate_core(ctx, nullptr, batch, pov_daa_score, flags, params, results);
}
extern "C" int kgv_validate_txs(kgv_ctx* ctx, kgv_utxo_table* t, const kgv_tx_batch* batch, uint64_t pov_daa_score, uint32_t flags, const kgv_params* params,
                                kgv_tx_result* results) {
  if (!ctx || !t) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  return validate_core(ctx, t, batch, pov_daa_score, flags, params, results);
}

extern "C" int kgv_utxo_apply_accepted(kgv_ctx* ctx, kgv_utxo_table* t, const kgv_tx_batch* batch, const uint8_t* accept, uint64_t pov_daa_score) {
  if (!ctx || !t) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!batch || (batch->n_txs && !accept)) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  if (batch->n_txs == 0) return KGV_OK;
  CK(cudaSetDevice(ctx->device));
  kgv_dev_batch d;
  int rc = kgv_batch_to_device(ctx, batch, &d, false);
  if (rc) return rc;
  size_t nt = d.n_txs, ni = d.n_inputs, no = d.n_outputs;
  size_t o_itx = 0, o_otx = al256(ni * 4), o_ids = al256(o_otx + no * 4), o_acc = al256(o_ids + nt * 32);
  rc = kgv_reserve(ctx, &ctx->d_scratch, &ctx->d_scratch_cap, al256(o_acc + nt));
  if (rc) return rc;
  uint8_t* S = ctx->d_scratch;
  const uint8_t* dacc = accept;
  if (!kgv_ptr_is_device(accept)) {
    CK(cudaMemcpyAsync(S + o_acc, accept, nt, cudaMemcpyHostToDevice, ctx->stream));
    dacc = S + o_acc;
  }
  BatchView v{d.txs, d.inputs, d.outputs, nullptr, d.bytes};
  cudaStream_t st = ctx->stream;
  k_tx_ids_dev<<<nblk(nt, 128), 128, 0, st>>>(v, (uint32_t)nt, (uint64_t*)(S + o_ids));
  CK(cudaGetLastError());
  k_input_tx_index<<<nblk(nt, 128), 128, 0, st>>>(d.txs, (uint32_t)nt, (uint32_t*)(S + o_itx));
  CK(cudaGetLastError());
  k_output_tx_index<<<nblk(nt, 128), 128, 0, st>>>(d.txs, (uint32_t)nt, (uint32_t*)(S + o_otx));
  CK(cudaGetLastError());
  ctx->launches += 3;
  if (ni) { k_apply_erase<<<nblk(ni, 128), 128, 0, st>>>(view_of(t), d.txs, d.inputs, ni, (const uint32_t*)(S + o_itx), dacc); CK(cudaGetLastError()); ctx->launches++; }
  if (no) { k_apply_insert<<<nblk(no, 128), 128, 0, st>>>(view_of(t), v, no, (const uint32_t*)(S + o_otx), dacc, (const uint64_t*)(S + o_ids), pov_daa_score); CK(cudaGetLastError()); ctx->launches++; }
  if (!kgv_ptr_is_device(accept)) CK(cudaStreamSynchronize(st));
  return KGV_OK;
}

// ---------------------------------------------------------------------------------------------
// K8 entry points
// ---------------------------------------------------------------------------------------------
extern "C" int kgv_muhash_txs(kgv_ctx* ctx, kgv_utxo_table* table, const kgv_tx_batch* batch, const uint8_t* accept, uint64_t pov_daa_score,
                              uint8_t* numerator384, uint8_t* denominator384) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!batch || !numerator384 || !denominator384 || (batch->n_txs && !accept)) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  if (kgv_ptr_is_device(numerator384) != kgv_ptr_is_device(denominator384)) { ctx->err = "outputs must both be host or both be device pointers"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  kgv_dev_batch d;
  d.n_txs = d.n_inputs = d.n_outputs = d.n_bytes = 0;
  if (batch->n_txs) {
    int rc = kgv_batch_to_device(ctx, batch, &d, table == nullptr);
    if (rc) return rc;
  }
  const size_t nt = d.n_txs, ni = d.n_inputs, no = d.n_outputs;
  uint32_t *e_den = nullptr, *e_num = nullptr;
  int rc = kgv_mu_reserve(ctx, ni, no, &e_den, &e_num);
  if (rc) return rc;
  if (nt) {
    size_t o_ent = 0, o_itx = al256(o_ent + ni * sizeof(DevEntry)), o_otx = al256(o_itx + ni * 4), o_ids = al256(o_otx + no * 4), o_acc = al256(o_ids + nt * 32);
    rc = kgv_reserve(ctx, &ctx->d_scratch, &ctx->d_scratch_cap, al256(o_acc + nt));
    if (rc) return rc;
    uint8_t* S = ctx->d_scratch;
    cudaStream_t st = ctx->stream;
    const uint8_t* dacc = accept;
    if (!kgv_ptr_is_device(accept)) {
      CK(cudaMemcpyAsync(S + o_acc, accept, nt, cudaMemcpyHostToDevice, st));
      dacc = S + o_acc;
    }
    DevEntry* dent = (DevEntry*)(S + o_ent);
    if (ni) {
      if (table) k_populate<<<nblk(ni, 128), 128, 0, st>>>(view_of(table), d.inputs, ni, dent);
      else k_entries_from_batch<<<nblk(ni, 128), 128, 0, st>>>(d.entries, d.bytes, ni, dent);
      CK(cudaGetLastError());
      ctx->launches++;
    }
    BatchView v{d.txs, d.inputs, d.outputs, dent, d.bytes};
    k_tx_ids_dev<<<nblk(nt, 128), 128, 0, st>>>(v, (uint32_t)nt, (uint64_t*)(S + o_ids));
    CK(cudaGetLastError());
    k_input_tx_index<<<nblk(nt, 128), 128, 0, st>>>(d.txs, (uint32_t)nt, (uint32_t*)(S + o_itx));
    CK(cudaGetLastError());
    k_output_tx_index<<<nblk(nt, 128), 128, 0, st>>>(d.txs, (uint32_t)nt, (uint32_t*)(S + o_otx));
    CK(cudaGetLastError());
    if (ni + no) {
      k_muhash_tx_elements<<<nblk(ni + no, 128), 128, 0, st>>>(v, ni, no, (const uint32_t*)(S + o_itx), (const uint32_t*)(S + o_otx), dacc,
                                                               (const uint64_t*)(S + o_ids), pov_daa_score, e_den, e_num);
      CK(cudaGetLastError());
    }
    ctx->launches += 4;
  }
  return kgv_mu_reduce(ctx, ni, no, numerator384, denominator384);
}

// MuHash of the whole UTXO set (the pruning-point / virtual UTXO commitment: MuHash::add_utxo for every entry,
// consensus/core/src/muhash.rs:28-33).  The table is walked in chunks; chunk products are multiplied in a last tree.
extern "C" int kgv_utxo_muhash(kgv_ctx* ctx, kgv_utxo_table* t, uint8_t* numerator384) {
  if (!ctx || !t) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!numerator384) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  if (t->base) { ctx->err = "count / digest / MuHash are defined on plain tables: commit the view first"; return KGV_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const uint64_t slots = t->mask + 1;
  const size_t chunk = slots < ((uint64_t)1 << 17) ? (size_t)slots : ((size_t)1 << 17);
  const size_t n_chunks = (size_t)((slots + chunk - 1) / chunk);
  int rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, (n_chunks + 1) * 384);
  if (rc) return rc;
  uint8_t* prods = ctx->d_out;                    // n_chunks x 384 B, contiguous values
  uint8_t* dummy_den = ctx->d_out + n_chunks * 384;
  for (size_t c = 0; c < n_chunks; c++) {
    const uint64_t first = (uint64_t)c * chunk;
    const size_t n = (size_t)((slots - first) < chunk ? (slots - first) : chunk);
    uint32_t *e_den = nullptr, *e_num = nullptr;
    rc = kgv_mu_reserve(ctx, 0, n, &e_den, &e_num);
    if (rc) return rc;
    k_muhash_table_elements<<<nblk(n, 128), 128, 0, ctx->stream>>>(view_of(t), first, n, e_num);
    CK(cudaGetLastError());
    ctx->launches++;
    rc = kgv_mu_reduce(ctx, 0, n, prods + 384 * c, dummy_den);
    if (rc) return rc;
  }
  if (n_chunks == 1) {
    const bool dev = kgv_ptr_is_device(numerator384);
    CK(cudaMemcpyAsync(numerator384, prods, 384, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    if (!dev) CK(cudaStreamSynchronize(ctx->stream));
    return KGV_OK;
  }
  uint32_t *e_den = nullptr, *e_num = nullptr;
  rc = kgv_mu_reserve(ctx, 0, n_chunks, &e_den, &e_num);
  if (rc) return rc;
  k_u3072_scatter<<<nblk(n_chunks, 128), 128, 0, ctx->stream>>>((const uint32_t*)prods, n_chunks, e_num);
  CK(cudaGetLastError());
  ctx->launches++;
  if (kgv_ptr_is_device(numerator384)) return kgv_mu_reduce(ctx, 0, n_chunks, numerator384, dummy_den);
  uint8_t den_host[384];
  return kgv_mu_reduce(ctx, 0, n_chunks, numerator384, den_host);
}

#include "kgv_replay_impl.cuh"
