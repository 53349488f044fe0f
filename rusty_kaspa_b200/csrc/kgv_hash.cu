// kgv_hash.cu — tx id / tx hash / sighash kernels and their C ABI entry points (include/kgv.h).
// One hash per thread; the tx records are read once, the digests written once: these kernels stream
// the batch (HBM-bound side of the path) — the ALU work is 12 BLAKE2b rounds per 128 B.
#include "kgv_internal.h"
#include "kgv_txhash.cuh"

#include <cstdio>

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

int kgv_batch_to_device(kgv_ctx* ctx, const kgv_tx_batch* b, kgv_dev_batch* out, bool need_entries) {
  if (!b) { ctx->err = "null batch"; return KGV_ERR_ARG; }
  if ((b->n_txs && !b->txs) || (b->n_inputs && !b->inputs) || (b->n_outputs && !b->outputs) || (b->n_bytes && !b->bytes) ||
      (need_entries && b->n_inputs && !b->entries)) {
    ctx->err = "batch array missing";
    return KGV_ERR_ARG;
  }
  out->n_txs = b->n_txs; out->n_inputs = b->n_inputs; out->n_outputs = b->n_outputs; out->n_bytes = b->n_bytes;
  const void* probe = b->n_txs ? (const void*)b->txs : (const void*)b->bytes;
  if (probe && kgv_ptr_is_device(probe)) {
    out->txs = b->txs; out->inputs = b->inputs; out->outputs = b->outputs; out->entries = b->entries; out->bytes = b->bytes;
    return KGV_OK;
  }
  size_t o_tx = 0;
  size_t o_in = al256(o_tx + b->n_txs * sizeof(kgv_tx));
  size_t o_out = al256(o_in + b->n_inputs * sizeof(kgv_input));
  size_t o_ent = al256(o_out + b->n_outputs * sizeof(kgv_output));
  size_t o_by = al256(o_ent + (b->entries ? b->n_inputs * sizeof(kgv_utxo_entry) : 0));
  size_t total = al256(o_by + b->n_bytes + 16);
  int rc = kgv_reserve(ctx, &ctx->d_batch, &ctx->d_batch_cap, total);
  if (rc) return rc;
  uint8_t* d = ctx->d_batch;
  if (b->n_txs) CK(cudaMemcpyAsync(d + o_tx, b->txs, b->n_txs * sizeof(kgv_tx), cudaMemcpyHostToDevice, ctx->stream));
  if (b->n_inputs) CK(cudaMemcpyAsync(d + o_in, b->inputs, b->n_inputs * sizeof(kgv_input), cudaMemcpyHostToDevice, ctx->stream));
  if (b->n_outputs) CK(cudaMemcpyAsync(d + o_out, b->outputs, b->n_outputs * sizeof(kgv_output), cudaMemcpyHostToDevice, ctx->stream));
  if (b->entries && b->n_inputs) CK(cudaMemcpyAsync(d + o_ent, b->entries, b->n_inputs * sizeof(kgv_utxo_entry), cudaMemcpyHostToDevice, ctx->stream));
  if (b->n_bytes) CK(cudaMemcpyAsync(d + o_by, b->bytes, b->n_bytes, cudaMemcpyHostToDevice, ctx->stream));
  out->txs = (const kgv_tx*)(d + o_tx);
  out->inputs = (const kgv_input*)(d + o_in);
  out->outputs = (const kgv_output*)(d + o_out);
  out->entries = b->entries ? (const kgv_utxo_entry*)(d + o_ent) : nullptr;
  out->bytes = d + o_by;
  return KGV_OK;
}

// ---------------------------------------------------------------------------------------------
template <bool HASH>
__global__ void __launch_bounds__(128) k_tx_digest(BatchView b, uint32_t n_txs, uint64_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_txs) return;
  uint64_t d[4];
  if (HASH) tx_hash(d, b, i); else tx_id(d, b, i);
#pragma unroll
  for (int k = 0; k < 4; k++) out[4 * (size_t)i + k] = d[k];
}

__global__ void __launch_bounds__(128) k_sighash_reused(BatchView b, uint32_t n_txs, SigHashReused* __restrict__ reused) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_txs) return;
  SigHashReused r;
  sighash_reused(r, b, i);
  reused[i] = r;
}

__global__ void k_entries_to_dev(const kgv_utxo_entry* __restrict__ in, const uint8_t* __restrict__ bytes, size_t n, DevEntry* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  kgv_utxo_entry e = in[i];
  DevEntry d;
  d.amount = e.amount; d.block_daa_score = e.block_daa_score; d.script = bytes + e.script_off; d.script_len = e.script_len;
  d.spk_version = e.spk_version; d.is_coinbase = e.is_coinbase; d.found = 1;
  out[i] = d;
}

__global__ void __launch_bounds__(128)
k_sighash_items(BatchView b, const SigHashReused* __restrict__ reused, const kgv_sighash_item* __restrict__ items, size_t n_items,
                uint32_t n_txs, uint32_t n_inputs, uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  kgv_sighash_item it = items[i];
  uint32_t w[8];
  if (it.tx >= n_txs || it.input >= n_inputs || !sighash_type_allowed(it.hash_type)) {
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = 0xFFFFFFFFu;
  } else {
    SigHashReused r = reused[it.tx];
    sighash_final(w, b, it.tx, it.input, it.hash_type, it.ecdsa != 0, r);
  }
#pragma unroll
  for (int k = 0; k < 8; k++) out[8 * i + k] = bswap32(w[k]);  // back to the digest's byte order
}

// ---------------------------------------------------------------------------------------------
static int digest_common(kgv_ctx* ctx, const kgv_tx_batch* batch, uint8_t* out32, bool hash) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!batch || (batch->n_txs && !out32)) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  if (batch->n_txs == 0) return KGV_OK;
  CK(cudaSetDevice(ctx->device));
  kgv_dev_batch d;
  int rc = kgv_batch_to_device(ctx, batch, &d, false);
  if (rc) return rc;
  bool out_dev = kgv_ptr_is_device(out32);
  uint8_t* dout = out32;
  if (!out_dev) {
    rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, d.n_txs * 32);
    if (rc) return rc;
    dout = ctx->d_out;
  }
  BatchView v{d.txs, d.inputs, d.outputs, nullptr, d.bytes};
  unsigned blocks = (unsigned)((d.n_txs + 127) / 128);
  if (hash) k_tx_digest<true><<<blocks, 128, 0, ctx->stream>>>(v, (uint32_t)d.n_txs, (uint64_t*)dout);
  else k_tx_digest<false><<<blocks, 128, 0, ctx->stream>>>(v, (uint32_t)d.n_txs, (uint64_t*)dout);
  CK(cudaGetLastError());
  ctx->launches++;
  if (!out_dev) {
    CK(cudaMemcpyAsync(out32, dout, d.n_txs * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return KGV_OK;
}
extern "C" int kgv_tx_ids(kgv_ctx* ctx, const kgv_tx_batch* batch, uint8_t* out32) { return digest_common(ctx, batch, out32, false); }
extern "C" int kgv_tx_hashes(kgv_ctx* ctx, const kgv_tx_batch* batch, uint8_t* out32) { return digest_common(ctx, batch, out32, true); }

extern "C" int kgv_sighash(kgv_ctx* ctx, const kgv_tx_batch* batch, const kgv_sighash_item* items, size_t n_items, uint8_t* out32) {
  if (!ctx) return KGV_ERR_ARG;
  std::lock_guard<std::mutex> g(ctx->mu);
  if (!batch || (n_items && (!items || !out32))) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  if (n_items == 0) return KGV_OK;
  CK(cudaSetDevice(ctx->device));
  kgv_dev_batch d;
  int rc = kgv_batch_to_device(ctx, batch, &d, true);
  if (rc) return rc;
  bool io_dev = kgv_ptr_is_device(items);
  if ((bool)kgv_ptr_is_device(out32) != io_dev) { ctx->err = "items and out32 must both be host or both be device pointers"; return KGV_ERR_ARG; }
  size_t o_reused = 0, o_ent = al256(d.n_txs * sizeof(SigHashReused)), o_items = al256(o_ent + d.n_inputs * sizeof(DevEntry));
  rc = kgv_reserve(ctx, &ctx->d_scratch, &ctx->d_scratch_cap, o_items + (io_dev ? 0 : n_items * sizeof(kgv_sighash_item)) + 256);
  if (rc) return rc;
  SigHashReused* dre = (SigHashReused*)(ctx->d_scratch + o_reused);
  const kgv_sighash_item* ditems = items;
  uint8_t* dout = out32;
  if (!io_dev) {
    CK(cudaMemcpyAsync(ctx->d_scratch + o_items, items, n_items * sizeof(kgv_sighash_item), cudaMemcpyHostToDevice, ctx->stream));
    ditems = (const kgv_sighash_item*)(ctx->d_scratch + o_items);
    rc = kgv_reserve(ctx, &ctx->d_out, &ctx->d_out_cap, n_items * 32);
    if (rc) return rc;
    dout = ctx->d_out;
  }
  DevEntry* dent = (DevEntry*)(ctx->d_scratch + o_ent);
  if (d.n_inputs) {
    k_entries_to_dev<<<(unsigned)((d.n_inputs + 127) / 128), 128, 0, ctx->stream>>>(d.entries, d.bytes, d.n_inputs, dent);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  BatchView v{d.txs, d.inputs, d.outputs, dent, d.bytes};
  k_sighash_reused<<<(unsigned)((d.n_txs + 127) / 128), 128, 0, ctx->stream>>>(v, (uint32_t)d.n_txs, dre);
  CK(cudaGetLastError());
  k_sighash_items<<<(unsigned)((n_items + 127) / 128), 128, 0, ctx->stream>>>(v, dre, ditems, n_items, (uint32_t)d.n_txs, (uint32_t)d.n_inputs, (uint32_t*)dout);
  CK(cudaGetLastError());
  ctx->launches += 2;
  if (!io_dev) {
    CK(cudaMemcpyAsync(out32, dout, n_items * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return KGV_OK;
}
