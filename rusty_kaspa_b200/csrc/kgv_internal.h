// kgv_internal.h — shared between the translation units of libkgv.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#ifndef KGV_BLOCK
#define KGV_BLOCK 128         // threads per block of the verification kernels
#endif
#ifndef KGV_BLOCKS_PER_SM
#define KGV_BLOCKS_PER_SM 3   // 3 x 64 KiB of per-thread tables in shared memory
#endif

#ifndef KGV_ITEMS
#define KGV_ITEMS 4           // signatures per thread sharing one modular inversion (see k_schnorr_verify)
#endif

struct kgv_dev_batch_fwd;
struct kgv_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t aux_stream = nullptr;              // fork/join side stream: ECDSA items verify beside the Schnorr items
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_time[3] = {};                    // kgv_replay_window: phase timing for kgv_replay_stats
  cudaEvent_t ev_chunk[32] = {};                  // upload-complete events of the chunked host-pointer verify path
  cudaStream_t stream = nullptr;
  uint32_t* gtab = nullptr;     // [2][65536][16] u32: v*G and v*2^128*G, affine
  uint8_t* d_in = nullptr;      // staging for host-pointer calls
  size_t d_in_cap = 0;
  uint8_t* d_out = nullptr;
  size_t d_out_cap = 0;
  uint8_t* d_batch = nullptr;   // staging for host-resident transaction batches
  size_t d_batch_cap = 0;
  // kgv_batch_prefetch: host batches uploaded ahead of their use on copy_stream.  Two slots, because the caller prefetches window i+1 BEFORE it
  // issues the (synchronous) call for window i, whose own prefetched copy is still waiting in the other slot.
  struct PrefetchSlot {
    uint8_t* buf = nullptr;
    size_t cap = 0;
    bool valid = false;
    const void *txs = nullptr, *inputs = nullptr, *outputs = nullptr, *entries = nullptr, *bytes = nullptr;
    size_t n_txs = 0, n_inputs = 0, n_outputs = 0, n_bytes = 0;
    cudaEvent_t done = nullptr;   // upload finished
    std::thread worker;           // range checks + upload run here, off the caller's critical path; joined by whoever consumes / reuses the slot
    int rc = 0;
    std::string err;
  } prefetch[2];
  int prefetch_next = 0;          // slot the next kgv_batch_prefetch overwrites when both are taken
  cudaEvent_t ev_prefetch = nullptr;
  cudaStream_t copy_stream = nullptr;  // its own stream: aux_stream carries the ECDSA half of a validation call
  uint8_t* d_scratch = nullptr; // per-call device scratch (sub-hashes, sig items, ...)
  size_t d_scratch_cap = 0;
  uint8_t* d_work = nullptr;    // per-call populated entries / input->tx index / verdicts of the validation calls
  size_t d_work_cap = 0;
  uint8_t* d_replay = nullptr;  // kgv_replay_window: window-wide state (tx ids, window map, script verdicts, accept mask)
  size_t d_replay_cap = 0;
  uint8_t* d_mu = nullptr;      // MuHash element arrays, product-tree levels and wide-product scratch rows
  size_t d_mu_cap = 0;
  // state of the last kgv_replay_window call, kept for kgv_replay_muhash (cleared by any call that stages another batch)
  struct {
    bool valid = false;
    kgv_dev_batch_fwd* unused_ = nullptr;
    const void *txs = nullptr, *inputs = nullptr, *outputs = nullptr, *bytes = nullptr;
    size_t nt = 0, ni = 0, no = 0, n_blocks = 0;
    size_t o_ids = 0, o_itx = 0, o_otx = 0, o_ent = 0, o_acc = 0, o_txb = 0, o_rng = 0;  // offsets into d_replay
  } last_replay;
  struct kgv_sigcache* sigcache = nullptr;  // kgv_set_sigcache: verdicts of the validation calls are looked up / remembered here
  struct kgv_comm* shard_comm = nullptr;  // kgv_set_sharding: signature checks of the validation calls are split over its ranks
  std::vector<uint8_t*> parked;  // outgrown per-call buffers, released when the caller synchronises / destroys the context (kgv_reserve)
  uint64_t launches = 0;
  int resident_blocks = 148 * KGV_BLOCKS_PER_SM;  // verification kernels: blocks that fit the device at once (persistent grid)
  std::recursive_mutex mu;  // recursive: the host-VM resolution inside a validation call re-enters the ABI (kgv_sighash, kgv_*_verify)
  std::string err;
};

int kgv_ptr_is_device(const void* p);
int kgv_reserve(kgv_ctx* ctx, uint8_t** buf, size_t* cap, size_t need);

// ---- transaction batches on the device (kgv_hash.cu) ----
#include "../../include/kgv.h"
struct kgv_dev_batch {
  const kgv_tx* txs;
  const kgv_input* inputs;
  const kgv_output* outputs;
  const kgv_utxo_entry* entries;
  const uint8_t* bytes;
  size_t n_txs, n_inputs, n_outputs, n_bytes;
};
// Makes the batch arrays device-resident (uploads host arrays into ctx staging; wraps device arrays).
int kgv_batch_to_device(kgv_ctx* ctx, const kgv_tx_batch* b, kgv_dev_batch* out, bool need_entries);

// Enqueue a verification kernel on device-resident SoA item arrays (no locking, no copies): used by the
// fused validation path.  ecdsa: pk stride 33, else 32.
int kgv_launch_verify(kgv_ctx* ctx, const uint8_t* dpk, const uint8_t* dmsg, const uint8_t* dsig, size_t n, uint8_t* dstatus, bool ecdsa,
                      cudaStream_t on = nullptr, bool use_on = false, const uint32_t* index = nullptr, const uint32_t* n_dev = nullptr);

// ---- MuHash product trees (kgv_muhash.cu) ----
// Reserve the level-0 element arrays of the two trees (denominator = removed elements, numerator = added elements):
// element e of a tree with n elements has its limb block i at E + (i*n + e)*8 words (kgv_u3072.cuh layout, stride n).
int kgv_mu_reserve(kgv_ctx* ctx, size_t n_den, size_t n_num, uint32_t** e_den, uint32_t** e_num);
// Multiply each tree down to one value (denominator on ctx->stream, numerator on the side stream) and write the two
// canonical residues (384 little-endian bytes each) to host or device memory.
int kgv_mu_reduce(kgv_ctx* ctx, size_t n_den, size_t n_num, uint8_t* out_num384, uint8_t* out_den384);
// out + g * out_pitch_words = product of the level-0 elements E[lo[g] .. hi[g]) with flags[flag_index[j]] != 0 (one 16-lane group per range)
int kgv_mu_range_products(kgv_ctx* ctx, const uint32_t* E, size_t stride, const uint8_t* flags, const uint32_t* flag_index, const uint32_t* lo, const uint32_t* hi, uint32_t n_segs,
                          uint32_t* out, size_t out_pitch_words, cudaStream_t st);
int kgv_mu_canonicalize(kgv_ctx* ctx, uint32_t* vals, size_t pitch_words, size_t n, cudaStream_t st);

// ---- shared pieces of the validation path (kgv_validate.cu) ----
struct kgv_utxo_table;

// ---- multi-GPU exchange used by the sharded script phase (kgv_comm.cu) ----
// Every rank contributes `per` bytes at buf + rank * per (device memory, n_ranks * per bytes in all); on return (stream order)
// buf holds all ranks' contributions.  Peer transport if the communicator is connected, else NCCL (in place).
int kgv_comm_exchange_slices(kgv_ctx* ctx, struct kgv_comm* c, uint8_t* buf, size_t per);
int kgv_comm_ranks(const struct kgv_comm* c, int* rank);

// ---- signature cache (kgv_sigcache.cu): device verdict table keyed by BLAKE2b-256(kind || sig || pk || msg) ----
// Looks every item up; status[i] = cached verdict or 0xFF; writes the digests (32 B per item, reused by the insert), the compacted list of
// misses and their count (device).  Enqueued on `st`.
int kgv_sigcache_lookup(kgv_ctx* ctx, struct kgv_sigcache* c, const uint8_t* pk, const uint8_t* msg, const uint8_t* sig, size_t n, bool ecdsa, uint8_t* status, uint8_t* digests,
                        uint32_t* miss_index, uint32_t* n_miss_dev, cudaStream_t st);
// remembers the verdicts (0 / 1 only, as the reference: parse errors never reach its cache) of the listed items
int kgv_sigcache_insert(kgv_ctx* ctx, struct kgv_sigcache* c, const uint8_t* status, const uint8_t* digests, const uint32_t* miss_index, const uint32_t* n_miss_dev, size_t n_max,
                        cudaStream_t st);
