// kgv_replay_impl.cuh — kgv_replay_window: the caller side of the hot path as ONE library call (included by kgv_validate.cu,
// whose kernels it reuses).
//
// Reference: VirtualStateProcessor::calculate_utxo_state / verify_expected_utxo_state
// (consensus/src/pipeline/virtual_processor/utxo_validation.rs:110-173,182-228): the blocks of a mergeset are processed
// SEQUENTIALLY against `selected_parent_utxo_view.compose(&ctx.mergeset_diff)`, the transactions of one block in parallel
// (:262-309), and every accepted transaction is folded into the diff (UtxoDiff::add_transaction, utxo_diff.rs:233-247)
// before the next block is looked at.  simpa prints the rate of exactly this loop (simpa/src/main.rs:454-460).
//
// A 10-BPS block carries a few hundred signatures - three orders of magnitude too few for a B200 - but signatures are
// context free given the spent output (the sighash reads only the entry's script_public_key and amount,
// sighash.rs:252-255, and both are fixed by the outpoint).  So the window is processed in two device-resident passes:
//
//   pre-check   every script of the window in ONE batch (plan/sighash/verify/resolve of kgv_validate.cu, millions of
//               signature checks per launch).  Spent outputs come from the UTXO table or, when the output is created inside
//               the window, from the creating transaction (found through a window hash map  tx id -> tx index  built on the
//               device).  Which of the two exists at the spending block's position is irrelevant for the script verdict.
//   in-order    one persistent single-CTA kernel walks the blocks: populate from the table (now position dependent),
//               UTXO-context rules, accept = context ok && scripts ok, erase spent / insert created entries, next block.
//               Three CTA barriers per block instead of >= 15 kernel launches; slots of the next block are prefetched
//               into L2 while the current block is decided.
//
// Result per transaction: the context verdict when the context rules fail, else the script verdict - the order
// validate_populated_transaction_and_get_fee reports them in (tx_validation_in_utxo_context.rs:34-61).

// ---------------------------------------------------------------------------------------------
// window map: tx id -> tx index (open addressing over indices; the ids array holds the keys)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t wm_hash(const uint64_t* id) { return id[0] ^ (id[1] * 0x9E3779B97F4A7C15ull) ^ (id[2] >> 17) ^ (id[3] << 13); }
__device__ __forceinline__ bool id_eq(const uint64_t* a, const uint64_t* b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3]; }

__global__ void k_wm_insert(const uint64_t* __restrict__ ids, uint32_t n_txs, uint32_t* __restrict__ wm, uint64_t wm_mask) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_txs) return;
  const uint64_t* id = ids + 4 * (size_t)t;
  uint64_t i = wm_hash(id) & wm_mask;
  for (uint64_t p = 0; p <= wm_mask; p++, i = (i + 1) & wm_mask) {
    uint32_t cur = atomicCAS(&wm[i], 0u, t + 1);
    if (cur == 0u) return;
    // the same transaction may sit in several parallel blocks of a DAG: one representative is enough (identical outputs)
    if (id_eq(ids + 4 * (size_t)(cur - 1), id)) return;
  }
}
__device__ __forceinline__ int wm_find(const uint64_t* __restrict__ ids, const uint32_t* __restrict__ wm, uint64_t wm_mask, const uint64_t* id) {
  uint64_t i = wm_hash(id) & wm_mask;
  for (uint64_t p = 0; p <= wm_mask; p++, i = (i + 1) & wm_mask) {
    uint32_t cur = wm[i];
    if (cur == 0u) return -1;
    if (id_eq(ids + 4 * (size_t)(cur - 1), id)) return (int)(cur - 1);
  }
  return -1;
}

// spent entry of every input for the pre-check: UTXO table first, else an output created inside the window
__global__ void k_populate_window(TableView t, BatchView b, size_t n_inputs, const uint64_t* __restrict__ ids, const uint32_t* __restrict__ wm, uint64_t wm_mask,
                                  DevEntry* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inputs) return;
  const kgv_input& in = b.inputs[i];
  uint32_t k[9];
  input_key(k, in);
  SlotHead h;
  UtxoSlot* s = table_find(t, k, h);
  DevEntry d;
  if (s) head_to_entry(d, t, s, h);
  else {
    entry_absent(d);
    uint64_t id[4];
#pragma unroll
    for (int w = 0; w < 4; w++) id[w] = (uint64_t)k[2 * w] | ((uint64_t)k[2 * w + 1] << 32);
    int src = wm_find(ids, wm, wm_mask, id);
    if (src >= 0) {
      const kgv_tx& stx = b.txs[src];
      if (in.prev_index < stx.n_outputs) {
        const kgv_output& o = b.outputs[stx.first_output + in.prev_index];
        d.amount = o.value; d.script = b.bytes + o.script_off; d.script_len = o.script_len; d.spk_version = o.spk_version;
        d.is_coinbase = tx_is_coinbase(stx) ? 1 : 0;
        d.found = 1;
      }
    }
  }
  out[i] = d;
}

// per-transaction block index and the pre-check's starting status
#define KGV_PRE_SKIPPED 0xFEu  // not script-checked in the pre-pass (coinbase position, SkipScriptChecks block)
__global__ void k_replay_tx_block(const kgv_replay_block* __restrict__ blocks, uint32_t n_blocks, uint32_t* __restrict__ tx_block) {
  uint32_t bi = blockIdx.x;
  if (bi >= n_blocks) return;
  kgv_replay_block bl = blocks[bi];
  for (uint32_t j = threadIdx.x; j < bl.n_txs; j += blockDim.x) tx_block[bl.first_tx + j] = bi;
}
__global__ void k_replay_pre_status(BatchView b, uint32_t n_txs, const kgv_replay_block* __restrict__ blocks, const uint32_t* __restrict__ tx_block,
                                    kgv_tx_result* __restrict__ pre) {
  uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
  if (ti >= n_txs) return;
  kgv_replay_block bl = blocks[tx_block[ti]];
  const kgv_tx& t = b.txs[ti];
  kgv_tx_result r;
  r.fee = 0; r.fail_input = 0; r.status = KGV_TX_OK; r.script_err = 0; r.pad_[0] = r.pad_[1] = 0;
  if (ti == bl.first_tx || tx_is_coinbase(t) || (bl.flags & KGV_REPLAY_SKIP_SCRIPTS)) r.status = KGV_PRE_SKIPPED;
  else {
    const DevEntry* ent = b.entries + t.first_input;
    for (uint32_t i = 0; i < t.n_inputs; i++)
      if (!ent[i].found) { r.status = KGV_TX_MISSING_OUTPOINTS; break; }  // cannot exist at its block's position either
  }
  pre[ti] = r;
}
// ---------------------------------------------------------------------------------------------
// the in-order pass: ONE CTA, blocks in sequence
// ---------------------------------------------------------------------------------------------
// per-block ranges, computed in parallel before the walk so that the in-order kernel never chases tx records to find them
struct ReplayRange {
  uint32_t t0, t1, i0, i1, o0, o1, flags, pad_;
  uint64_t pov;
};
__global__ void k_replay_ranges(const kgv_replay_block* __restrict__ blocks, uint32_t n_blocks, const kgv_tx* __restrict__ txs, ReplayRange* __restrict__ out) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  kgv_replay_block bl = blocks[b];
  ReplayRange r;
  r.t0 = bl.first_tx; r.t1 = bl.first_tx + bl.n_txs; r.flags = bl.flags; r.pad_ = 0; r.pov = bl.pov_daa_score;
  r.i0 = r.i1 = r.o0 = r.o1 = 0;
  if (bl.n_txs) {
    const kgv_tx& tf = txs[r.t0];
    const kgv_tx& tl = txs[r.t1 - 1];
    r.i0 = tf.first_input; r.i1 = tl.first_input + tl.n_inputs; r.o0 = tf.first_output; r.o1 = tl.first_output + tl.n_outputs;
  }
  out[b] = r;
}

struct ReplayArgs {
  TableView t;
  BatchView b;              // b.entries = dent (written here: the entry every input finds AT ITS BLOCK'S POSITION)
  DevEntry* dent;
  uint8_t* spent_scripts;   // 72 bytes per input: copy of an inline script (the slot may be reused later in the window)
  UtxoSlot** slotp;         // slot of every input (for the erase)
  const uint64_t* ids;
  const uint32_t* itx;
  const uint32_t* otx;
  const ReplayRange* ranges;
  uint32_t n_blocks;
  kgv_params prm;
  const kgv_tx_result* pre; // script verdicts of the pre-check
  kgv_tx_result* res;       // final verdicts
  uint8_t* accept;
  unsigned long long* stats; // [0] accepted transactions
};

// L2 prefetch of the 128-byte lines covering [p, p + bytes), dealt to the threads from the TOP of the CTA downwards (the low
// threads carry the per-input / per-transaction work of the current block)
__device__ __forceinline__ void prefetch_range(const void* p, size_t bytes, uint32_t rtid, uint32_t nth) {
  if (!bytes) return;
  const uintptr_t lo = (uintptr_t)p & ~(uintptr_t)127, hi = (uintptr_t)p + bytes;
  for (uintptr_t q = lo + 128 * (uintptr_t)rtid; q < hi; q += 128 * (uintptr_t)nth) prefetch_l2((const void*)q);
}

// The walk is a chain of dependent memory accesses per block (record -> key -> slot -> entry -> verdict -> slot update); left to
// itself every link is a DRAM miss (~1 us) and a block costs ~20 us.  So the kernel runs a two-block look-ahead entirely with
// L2 prefetches: while block b is decided, the records of block b+2 (transactions, inputs, outputs, ids, script verdicts) are
// prefetched by address range, and for block b+1 - whose records are L2 hits by then - the table slots its inputs will probe,
// the slots its outputs will be inserted into and the script bytes those inserts copy.  The walk itself then only meets L2 hits.
__global__ void __launch_bounds__(1024, 1) k_replay_inorder(ReplayArgs a) {
  const uint32_t tid = threadIdx.x, nth = blockDim.x, rtid = nth - 1 - tid;
  __shared__ unsigned long long s_acc;
  __shared__ int s_live, s_tomb;  // table counter deltas of this launch (one global atomic at the end instead of one per entry)
  if (tid == 0) { s_acc = 0; s_live = 0; s_tomb = 0; }
  auto prefetch_records = [&](uint32_t bi) {
    if (bi >= a.n_blocks) return;
    const ReplayRange r = a.ranges[bi];
    prefetch_range(a.b.txs + r.t0, (size_t)(r.t1 - r.t0) * sizeof(kgv_tx), rtid, nth);
    prefetch_range(a.b.inputs + r.i0, (size_t)(r.i1 - r.i0) * sizeof(kgv_input), rtid, nth);
    prefetch_range(a.b.outputs + r.o0, (size_t)(r.o1 - r.o0) * sizeof(kgv_output), rtid, nth);
    prefetch_range(a.ids + 4 * (size_t)r.t0, (size_t)(r.t1 - r.t0) * 32, rtid, nth);
    prefetch_range(a.pre + r.t0, (size_t)(r.t1 - r.t0) * sizeof(kgv_tx_result), rtid, nth);
    prefetch_range(a.itx + r.i0, (size_t)(r.i1 - r.i0) * 4, rtid, nth);
    prefetch_range(a.otx + r.o0, (size_t)(r.o1 - r.o0) * 4, rtid, nth);
  };
  auto prefetch_slots = [&](uint32_t bi) {
    if (bi >= a.n_blocks) return;
    const ReplayRange r = a.ranges[bi];
    for (uint32_t i = r.i0 + rtid; i < r.i1; i += nth) {
      uint32_t k[9];
      input_key(k, a.b.inputs[i]);
      prefetch_l2(&a.t.slots[key_hash(k) & a.t.mask]);
    }
    if (!(r.flags & KGV_REPLAY_VERIFY_ONLY))
      for (uint32_t o = r.o0 + rtid; o < r.o1; o += nth) {
        const uint32_t ti = a.otx[o];
        uint32_t k[9];
#pragma unroll
        for (int w = 0; w < 4; w++) { uint64_t q = a.ids[4 * (size_t)ti + w]; k[2 * w] = (uint32_t)q; k[2 * w + 1] = (uint32_t)(q >> 32); }
        k[8] = o - a.b.txs[ti].first_output;
        prefetch_l2(&a.t.slots[key_hash(k) & a.t.mask]);
        prefetch_l2(a.b.bytes + a.b.outputs[o].script_off);
      }
  };
  prefetch_records(0);
  prefetch_records(1);
  __syncthreads();
  prefetch_slots(0);
  for (uint32_t bi = 0; bi < a.n_blocks; bi++) {
    const ReplayRange bl = a.ranges[bi];
    prefetch_records(bi + 2);
    if (bl.t1 == bl.t0) { prefetch_slots(bi + 1); continue; }
    const uint32_t i0 = bl.i0, i1 = bl.i1, o0 = bl.o0, o1 = bl.o1;
    // ---- A: populate from the table as it stands after the previous block (utxo_validation.rs:319-327)
    for (uint32_t i = i0 + tid; i < i1; i += nth) {
      uint32_t k[9];
      input_key(k, a.b.inputs[i]);
      SlotHead h;
      UtxoSlot* s = table_find(a.t, k, h);
      DevEntry d;
      if (s) {
        head_to_entry(d, a.t, s, h);
        if (d.script_len <= INLINE_SCRIPT) {  // keep the bytes: MuHash / diff consumers read them after the slot may have been reused
          uint32_t* dst = (uint32_t*)(a.spent_scripts + 72 * (size_t)i);
          const uint32_t* src = (const uint32_t*)((const uint8_t*)s + 64);
          dst[0] = h.w[15];
          const uint32_t nw = (d.script_len + 3) >> 2;
          for (uint32_t w = 1; w < nw; w++) dst[w] = __ldcg(src + (w - 1));
          d.script = (const uint8_t*)dst;
        }
      } else entry_absent(d);
      a.dent[i] = d;
      a.slotp[i] = s;
    }
    prefetch_slots(bi + 1);
    __syncthreads();
    // ---- B: context rules and the acceptance decision
    for (uint32_t ti = bl.t0 + tid; ti < bl.t1; ti += nth) {
      const bool cb = ti == bl.t0 || tx_is_coinbase(a.b.txs[ti]);
      kgv_tx_result r = tx_context_rules(a.b, ti, bl.pov, KGV_FLAGS_SKIP_SCRIPT_CHECKS, a.prm, cb);
      bool acc;
      if (cb) acc = (ti == bl.t0) && (bl.flags & KGV_REPLAY_ACCEPT_COINBASE);
      else {
        acc = r.status == KGV_TX_OK;
        if (acc && !(bl.flags & KGV_REPLAY_SKIP_SCRIPTS)) {
          const kgv_tx_result p = a.pre[ti];
          if (p.status != KGV_TX_OK) { r.status = p.status; r.script_err = p.script_err; r.fail_input = p.fail_input; acc = false; }
        }
      }
      if (bl.flags & KGV_REPLAY_VERIFY_ONLY) acc = false;
      a.res[ti] = r;
      a.accept[ti] = acc ? 1 : 0;
      if (acc && !cb) atomicAdd(&s_acc, 1ull);
    }
    __syncthreads();
    // ---- C: UtxoDiff::add_transaction straight into the table (utxo_diff.rs:233-247).  One CTA: the barrier orders these writes
    // before the next block's probes, no device-wide fence is needed inside the walk.
    if (!(bl.flags & KGV_REPLAY_VERIFY_ONLY)) {
      for (uint32_t i = i0 + tid; i < i1; i += nth) {
        if (!a.accept[a.itx[i]]) continue;
        UtxoSlot* s = a.slotp[i];
        *(volatile uint32_t*)&s->state = SLOT_TOMB;
        atomicSub(&s_live, 1);
        atomicAdd(&s_tomb, 1);
      }
      for (uint32_t o = o0 + tid; o < o1; o += nth) {
        const uint32_t ti = a.otx[o];
        if (!a.accept[ti]) continue;
        const kgv_tx& tx = a.b.txs[ti];
        const kgv_output& out = a.b.outputs[o];
        uint32_t k[9];
#pragma unroll
        for (int w = 0; w < 4; w++) { uint64_t q = a.ids[4 * (size_t)ti + w]; k[2 * w] = (uint32_t)q; k[2 * w + 1] = (uint32_t)(q >> 32); }
        k[8] = o - tx.first_output;
        table_put<false>(a.t, k, out.value, bl.pov, out.spk_version, (ti == bl.t0 || tx_is_coinbase(tx)) ? 1u : 0u, a.b.bytes + out.script_off, out.script_len, &s_live, &s_tomb);
      }
      __syncthreads();
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    a.stats[0] = s_acc;
    atomicAdd(&a.t.counters[0], (unsigned long long)(long long)s_live);
    atomicAdd(&a.t.counters[1], (unsigned long long)(long long)s_tomb);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
extern "C" int kgv_replay_window(kgv_ctx* ctx, kgv_utxo_table* table, const kgv_tx_batch* batch, const kgv_replay_block* blocks, size_t n_blocks,
                                 const kgv_params* prm, kgv_tx_result* results, uint8_t* accept, kgv_replay_stats* stats) {
  if (!ctx || !table) return KGV_ERR_ARG;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!batch || !prm || (n_blocks && !blocks) || (batch->n_txs && !results)) { ctx->err = "null argument"; return KGV_ERR_ARG; }
  if (stats) { stats->n_accepted = 0; stats->n_sig_checks = 0; stats->n_host_vm = 0; stats->pre_check_ms = 0; stats->in_order_ms = 0; }
  if (batch->n_txs == 0 || n_blocks == 0) return KGV_OK;
  if (n_blocks > 0xFFFFFFFFull) { ctx->err = "too many blocks"; return KGV_ERR_ARG; }
  // the blocks must tile the batch in order (block b = transactions [first_tx, first_tx + n_txs))
  {
    uint64_t at = 0;
    for (size_t i = 0; i < n_blocks; i++) {
      if (blocks[i].first_tx != at || blocks[i].flags > 7u) { ctx->err = "replay blocks must tile the batch contiguously, in order, with known flags"; return KGV_ERR_ARG; }
      at += blocks[i].n_txs;
    }
    if (at != batch->n_txs) { ctx->err = "replay blocks do not cover the batch"; return KGV_ERR_ARG; }
  }
  CK(cudaSetDevice(ctx->device));
  kgv_dev_batch d;
  int rc = kgv_batch_to_device(ctx, batch, &d, false);
  if (rc) return rc;
  const size_t nt = d.n_txs, ni = d.n_inputs, no = d.n_outputs;
  uint64_t wm_cap = 1024;
  while (wm_cap < 2 * nt) wm_cap <<= 1;
  // window state
  size_t o_ids = 0;
  size_t o_wm = al256(o_ids + nt * 32);
  size_t o_blk = al256(o_wm + wm_cap * 4);
  size_t o_txb = al256(o_blk + n_blocks * sizeof(kgv_replay_block));
  size_t o_itx = al256(o_txb + nt * 4);
  size_t o_otx = al256(o_itx + ni * 4);
  size_t o_ent = al256(o_otx + no * 4);
  size_t o_pre = al256(o_ent + ni * sizeof(DevEntry));
  size_t o_res = al256(o_pre + nt * sizeof(kgv_tx_result));
  size_t o_acc = al256(o_res + nt * sizeof(kgv_tx_result));
  size_t o_scr = al256(o_acc + nt);
  size_t o_slp = al256(o_scr + ni * 72);
  size_t o_rng = al256(o_slp + ni * sizeof(UtxoSlot*));
  size_t o_cnt = al256(o_rng + n_blocks * sizeof(ReplayRange));
  size_t total = al256(o_cnt + 64);
  rc = kgv_reserve(ctx, &ctx->d_replay, &ctx->d_replay_cap, total);
  if (rc) return rc;
  uint8_t* R = ctx->d_replay;
  uint64_t* ids = (uint64_t*)(R + o_ids);
  uint32_t* wm = (uint32_t*)(R + o_wm);
  kgv_replay_block* dblk = (kgv_replay_block*)(R + o_blk);
  uint32_t *txb = (uint32_t*)(R + o_txb), *itx = (uint32_t*)(R + o_itx), *otx = (uint32_t*)(R + o_otx);
  DevEntry* dent = (DevEntry*)(R + o_ent);
  kgv_tx_result *pre = (kgv_tx_result*)(R + o_pre), *res = (kgv_tx_result*)(R + o_res);
  uint8_t* dacc = R + o_acc;
  unsigned long long* cnt = (unsigned long long*)(R + o_cnt);
  cudaStream_t st = ctx->stream;
  if (stats) {
    for (cudaEvent_t& e : ctx->ev_time) if (!e) CK(cudaEventCreate(&e));
    CK(cudaEventRecord(ctx->ev_time[0], st));
  }
  CK(cudaMemcpyAsync(dblk, blocks, n_blocks * sizeof(kgv_replay_block), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(wm, 0, wm_cap * 4, st));
  CK(cudaMemsetAsync(cnt, 0, 64, st));
  BatchView v0{d.txs, d.inputs, d.outputs, nullptr, d.bytes};
  k_tx_ids_dev<<<nblk(nt, 128), 128, 0, st>>>(v0, (uint32_t)nt, ids);
  CK(cudaGetLastError());
  k_wm_insert<<<nblk(nt, 128), 128, 0, st>>>(ids, (uint32_t)nt, wm, wm_cap - 1);
  CK(cudaGetLastError());
  k_input_tx_index<<<nblk(nt, 128), 128, 0, st>>>(d.txs, (uint32_t)nt, itx);
  CK(cudaGetLastError());
  k_output_tx_index<<<nblk(nt, 128), 128, 0, st>>>(d.txs, (uint32_t)nt, otx);
  CK(cudaGetLastError());
  k_replay_tx_block<<<(unsigned)n_blocks, 128, 0, st>>>(dblk, (uint32_t)n_blocks, txb);
  CK(cudaGetLastError());
  k_replay_ranges<<<nblk(n_blocks, 128), 128, 0, st>>>(dblk, (uint32_t)n_blocks, d.txs, (ReplayRange*)(R + o_rng));
  CK(cudaGetLastError());
  ctx->launches += 6;
  // ---- pre-check of every script of the window
  BatchView v{d.txs, d.inputs, d.outputs, dent, d.bytes};
  if (ni) {
    k_populate_window<<<nblk(ni, 128), 128, 0, st>>>(view_of(table), v, ni, ids, wm, wm_cap - 1, dent);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  k_replay_pre_status<<<nblk(nt, 128), 128, 0, st>>>(v, (uint32_t)nt, dblk, txb, pre);
  CK(cudaGetLastError());
  ctx->launches++;
  STAGE("replay pre-status");
  uint64_t n_items = 0;
  rc = kgv_scripts_phase(ctx, v, nt, ni, itx, pre, &n_items);
  if (rc) return rc;
  k_count_status<<<nblk(nt, 256), 256, 0, st>>>(pre, (uint32_t)nt, KGV_TX_NEEDS_HOST_VM, cnt + 1);
  CK(cudaGetLastError());
  ctx->launches++;
  unsigned long long n_vm = 0;
  CK(cudaMemcpyAsync(&n_vm, cnt + 1, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (n_vm) {
    // non-standard scripts: the host engine decides them on the entries the pre-check populated (kgv_host_vm_resolve)
    rc = kgv_host_vm_resolve(ctx, batch, d, dent, pre);
    if (rc) return rc;
    // the host engine's GPU rounds went through the batch staging buffer: stage the window again
    rc = kgv_batch_to_device(ctx, batch, &d, false);
    if (rc) return rc;
    v = BatchView{d.txs, d.inputs, d.outputs, dent, d.bytes};
  }
  // ---- in-order pass
  ReplayArgs a;
  a.t = view_of(table);
  a.b = v;
  a.dent = dent;
  a.spent_scripts = R + o_scr;
  a.slotp = (UtxoSlot**)(R + o_slp);
  a.ids = ids; a.itx = itx; a.otx = otx;
  a.ranges = (const ReplayRange*)(R + o_rng); a.n_blocks = (uint32_t)n_blocks;
  a.prm = *prm;
  a.pre = pre; a.res = res; a.accept = dacc; a.stats = cnt;
  if (stats) CK(cudaEventRecord(ctx->ev_time[1], st));
  k_replay_inorder<<<1, 1024, 0, st>>>(a);
  CK(cudaGetLastError());
  ctx->launches++;
  if (stats) CK(cudaEventRecord(ctx->ev_time[2], st));
  STAGE("replay in-order");
  const bool dev_out = kgv_ptr_is_device(results);
  CK(cudaMemcpyAsync(results, res, nt * sizeof(kgv_tx_result), dev_out ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  if (accept) CK(cudaMemcpyAsync(accept, dacc, nt, kgv_ptr_is_device(accept) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  unsigned long long n_acc = 0;
  if (stats || !dev_out) {
    CK(cudaMemcpyAsync(&n_acc, cnt, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  if (stats) {
    stats->n_accepted = n_acc; stats->n_sig_checks = n_items; stats->n_host_vm = n_vm;
    CK(cudaEventElapsedTime(&stats->pre_check_ms, ctx->ev_time[0], ctx->ev_time[1]));
    CK(cudaEventElapsedTime(&stats->in_order_ms, ctx->ev_time[1], ctx->ev_time[2]));
  }
  return KGV_OK;
}
