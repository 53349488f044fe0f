/*
 * kgv.h — C ABI of the B200 transaction-validation library (libkgv.so).
 *
 * "kgv" = Kaspa GPU Validator.  This is the drop-in boundary for the hot path named by
 * BASELINE.json: batched secp256k1 Schnorr/ECDSA verification, sighash / tx-id hashing, the
 * UTXO table and the fused per-transaction validation that rusty-kaspa fans out over rayon in
 *   consensus/src/pipeline/virtual_processor/utxo_validation.rs:262-309
 *   consensus/src/processes/transaction_validator/tx_validation_in_utxo_context.rs:34-61,157-196
 *   crypto/txscript/src/lib.rs:574-643
 * The reference has no FFI layer for this path: these entry points are what a
 * Rust shim (`extern "C"` block, shown in INTEGRATION.md) binds to stand behind the reference's
 * own TransactionValidator / SigCache / UtxoView surface.
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every buffer; the library never keeps a host
 *     pointer past the call.
 *   - every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer on the
 *     context's device; the library detects which (cudaPointerGetAttributes).  All data
 *     pointers of one call must be of the same kind.  With host pointers the call copies in,
 *     computes, copies out and returns after the results are in the caller's memory.  With
 *     device pointers the work is enqueued on the context's stream and the call returns
 *     without synchronising (use kgv_synchronize or your own stream sync).
 *   - return value: 0 = ok, negative = argument / CUDA / NCCL failure (kgv_last_error explains).
 *     An invalid signature is NEVER an error return: verdicts are per-item status bytes.
 *   - there is no CPU fallback: without a usable CUDA device kgv_create fails.
 *   - results are deterministic and independent of batch split or GPU count.
 *   - a context serialises its calls with an internal mutex; use one context per thread for
 *     concurrency.
 */
#ifndef KGV_H
#define KGV_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGV_OK 0
#define KGV_ERR_ARG (-1)
#define KGV_ERR_CUDA (-2)
#define KGV_ERR_NOMEM (-3)
#define KGV_ERR_NCCL (-4)

/* Per-signature verdicts.  The reference distinguishes these cases
 * (crypto/txscript/src/lib.rs:582-583, 593, 618-619, 628):
 *   a malformed key / overflowing ECDSA r|s aborts the script with InvalidSignature,
 *   a well-formed but wrong signature is Ok(false) and (in multisig) the loop moves on. */
#define KGV_SIG_INVALID 0       /* sig.verify(..) -> Err  => Ok(false)                       */
#define KGV_SIG_VALID 1         /* sig.verify(..) -> Ok   => Ok(true)                        */
#define KGV_SIG_PK_PARSE_ERR 2  /* XOnlyPublicKey::from_slice / PublicKey::from_slice failed */
#define KGV_SIG_SIG_PARSE_ERR 3 /* ecdsa::Signature::from_compact failed (r or s >= n)       */

typedef struct kgv_ctx kgv_ctx;

/* One context per device.  Builds the generator window tables (8 MiB, L2 resident) on the GPU. */
int kgv_create(int device, uint32_t flags, kgv_ctx** out);
void kgv_destroy(kgv_ctx* ctx);
/* Run all subsequent work of this context on the given cudaStream_t (NULL = CUDA's default
 * stream, as for any cudaStream_t).  Lets a caller bracket kernels with its own events (bench.py
 * passes torch's stream).  kgv_reset_stream goes back to the context's private stream. */
int kgv_set_stream(kgv_ctx* ctx, void* cuda_stream);
int kgv_reset_stream(kgv_ctx* ctx);
int kgv_synchronize(kgv_ctx* ctx);
const char* kgv_last_error(const kgv_ctx* ctx);
/* Number of kernel launches this context has issued so far (bench.py's gpu_launches). */
uint64_t kgv_launch_count(const kgv_ctx* ctx);

/* Batched BIP-340 Schnorr verification: status[i] = verdict of (pk32[i], msg32[i], sig64[i]).
 * Replaces the per-signature FFI call `sig.verify(&msg,&pk)` of
 * crypto/txscript/src/lib.rs:593 together with the parses at :582-583.
 * SoA layout: pk32 = n*32 bytes (x-only key), msg32 = n*32 bytes, sig64 = n*64 bytes (r||s). */
int kgv_schnorr_verify(kgv_ctx* ctx, const uint8_t* pk32, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status);

/* Batched ECDSA verification (compressed 33-byte keys, 64-byte compact signatures), with
 * libsecp256k1 semantics: low-S required, r|s >= n is a parse error.
 * Replaces crypto/txscript/src/lib.rs:618-628. */
int kgv_ecdsa_verify(kgv_ctx* ctx, const uint8_t* pk33, const uint8_t* msg32, const uint8_t* sig64, size_t n, uint8_t* status);

/* Pack verdicts into a validity bitmap: bit i (LSB-first within each byte) = (status[i] == KGV_SIG_VALID).
 * bitmap has (n+7)/8 bytes.  This is the per-shard payload of the multi-GPU all-gather. */
int kgv_status_to_bitmap(kgv_ctx* ctx, const uint8_t* status, size_t n, uint8_t* bitmap);

/* Test / audit hook: affine coordinates (x||y, 32-byte big-endian each) of entry v (1..65535) of
 * generator table `which` (0: v*G, 1: v*2^128*G) as built on the device. */
int kgv_gtable_entry(kgv_ctx* ctx, int which, uint32_t v, uint8_t out_xy[64]);

/* Test / audit hook: verifies ONE Schnorr triple (host pointers) on the device and returns the
 * traced intermediates: trace_words[stage*16 + i], KGV_TRACE_STAGES stages of 16 u32 words
 * (stage numbering in csrc/kgv_verify.cuh).  tests/ compare it with the host build of the same
 * device code to localise a divergence. */
#define KGV_TRACE_STAGES 32
int kgv_debug_schnorr_trace(kgv_ctx* ctx, const uint8_t* pk32, const uint8_t* msg32, const uint8_t* sig64, uint32_t* trace_words,
                            uint8_t* status);

#ifdef __cplusplus
}
#endif
#endif /* KGV_H */
