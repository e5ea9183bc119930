// Multi-GPU exchange for sharded scans: every rank evaluates the rows of its
// shard against the same center; what has to cross GPUs per scan is one 32-byte summary (count,
// positives, first-max arg-max).  The reference's counterpart is the OpenMP reduction at the end of
// Trainer::get_close (custom `pmax` arg-max + `&&`, Trainer.cpp:38-48,81).
//
// The scan kernel itself stores each CTA's partial into every rank's inbox over NVLink peer memory
// (McPeerPush in mc_common.cuh, {data, epoch} words); scan_combine_kernel below waits for the
// world x num_sms records of a slot and folds them with the reference's rule (largest f0, first
// row among equals).  No host round trip, no collective library call, nothing but posted stores on
// the critical path.  Rows are numbered globally on every rank (the histogram matrix is replicated
// once after K1; only scan work and alive flags are sharded, in blocks of ~256 KB of consecutive rows
// dealt round-robin to the ranks), so records compare directly.
#include <string.h>

#include "mc_common.cuh"

int mc_launch_scan_push(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked,
                        void *partials_dev, int *nparts_out, const McPeerPush *push);

int mc_comm_scan_push(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked, int slot, int fence);

struct CombineArgs {
	unsigned long long slot_off[MC_XSLOTS];
	unsigned int epoch[MC_XSLOTS];
};

constexpr int COMBINE_THREADS = 256;
constexpr unsigned long long COMBINE_TIMEOUT_NS = 4000000000ull;   // a peer that never sends is an error, not a hang

__device__ __forceinline__ uint4 ld_volatile16(const void *p) {
	uint4 r;
	asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
	return r;
}

// one CTA per slot
__global__ void __launch_bounds__(COMBINE_THREADS)
scan_combine_kernel(const uint8_t *__restrict__ inbox, int world, int nparts, CombineArgs args, int slot0,
                    mc_scan_result *__restrict__ out, unsigned int *__restrict__ err) {
	__shared__ mc_scan_result s_part[COMBINE_THREADS / 32];
	const int slot = slot0 + blockIdx.x;
	const unsigned int epoch = args.epoch[slot];
	const uint8_t *base = inbox + args.slot_off[slot];
	mc_scan_result mine;
	mine.n_eval = 0; mine.n_pos = 0; mine.best_row = -1; mine.best_f0 = -1.0;
	unsigned long long t0;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	bool failed = false;
	for (int i = threadIdx.x; i < world * nparts; i += blockDim.x) {
		const int rank = i / nparts, cta = i % nparts;
		const uint8_t *rec = base + ((size_t)rank * MC_SCAN_PARTS + cta) * MC_LL_RECORD_BYTES;
		uint4 q[4];
		for (;;) {
#pragma unroll
			for (int j = 0; j < 4; j++) q[j] = ld_volatile16(rec + j * 16);
			bool ok = true;
#pragma unroll
			for (int j = 0; j < 4; j++) ok = ok && q[j].y == epoch && q[j].w == epoch;
			if (ok) break;
			__nanosleep(64);   // the scans of the next burst may share this SM
			unsigned long long t1;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
			if (t1 - t0 > COMBINE_TIMEOUT_NS) { failed = true; break; }
		}
		if (failed) break;
		mc_scan_result r;
		r.n_eval = (long long)((unsigned long long)q[0].x | ((unsigned long long)q[0].z << 32));
		r.n_pos = (long long)((unsigned long long)q[1].x | ((unsigned long long)q[1].z << 32));
		r.best_row = (long long)((unsigned long long)q[2].x | ((unsigned long long)q[2].z << 32));
		r.best_f0 = __longlong_as_double((long long)((unsigned long long)q[3].x | ((unsigned long long)q[3].z << 32)));
		mc_scan_merge(mine, r);
	}
	if (failed) atomicExch(err, 1u);
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	mc_scan_warp_fold(mine);
	if (lane == 0) s_part[wib] = mine;
	__syncthreads();
	if (threadIdx.x == 0) {
		mc_scan_result r = s_part[0];
		for (int w = 1; w < (int)(blockDim.x >> 5); w++) mc_scan_merge(r, s_part[w]);
		out[slot] = r;
	}
}

// ---------------------------------------------------------------------------------------------
static void comm_release(mc_ctx *ctx) {
	McComm &cm = ctx->comm;
	for (int p = 0; p < MC_MAX_PEERS; p++) {
		if (cm.ipc_opened[p] && cm.peer_inbox[p]) cudaIpcCloseMemHandle(cm.peer_inbox[p]);
		cm.peer_inbox[p] = nullptr;
		cm.ipc_opened[p] = false;
	}
	cudaFree(cm.inbox);
	cudaFree(cm.d_out);
	if (cm.h_out) { cudaFreeHost(cm.h_out); cudaEventDestroy(cm.done); }
	if (cm.xstream) { cudaStreamSynchronize(cm.xstream); cudaStreamDestroy(cm.xstream); }
	for (int b = 0; b < 4; b++) if (cm.burst_done[b]) cudaEventDestroy(cm.burst_done[b]);
	cudaFree(cm.d_ll_partials);
	cm = McComm();
}

void mc_comm_destroy(mc_ctx *ctx) {
	if (ctx && ctx->comm.world) comm_release(ctx);
}

extern "C" int mc_comm_init(mc_ctx *ctx, int rank, int world, uint8_t *handle_out) {
	MC_REQUIRE(ctx, MC_ERR_ARG, "ctx is NULL");
	MC_REQUIRE(world >= 1 && world <= MC_MAX_PEERS && rank >= 0 && rank < world, MC_ERR_ARG, "rank %d / world %d invalid (at most %d ranks)", rank, world, MC_MAX_PEERS);
	MC_CUDA(cudaSetDevice(ctx->device));
	MC_CUDA(cudaStreamSynchronize(ctx->stream));
	if (ctx->comm.world) comm_release(ctx);
	McComm &cm = ctx->comm;
	MC_CUDA(cudaMalloc(&cm.inbox, MC_INBOX_BYTES));
	MC_CUDA(cudaMemset(cm.inbox, 0, MC_INBOX_BYTES));
	MC_CUDA(cudaMalloc(&cm.d_out, (size_t)MC_XSLOTS * sizeof(mc_scan_result) + 64));
	MC_CUDA(cudaMemset(cm.d_out, 0, (size_t)MC_XSLOTS * sizeof(mc_scan_result) + 64));
	cm.world = world;
	cm.rank = rank;
	cm.peer_inbox[rank] = cm.inbox;
	cm.connected = world == 1;
	if (handle_out) {
		static_assert(sizeof(cudaIpcMemHandle_t) <= MC_COMM_HANDLE_BYTES, "handle size");
		cudaIpcMemHandle_t h;
		memset(handle_out, 0, MC_COMM_HANDLE_BYTES);
		MC_CUDA(cudaIpcGetMemHandle(&h, cm.inbox));
		memcpy(handle_out, &h, sizeof(h));
	}
	return MC_OK;
}

extern "C" int mc_comm_connect(mc_ctx *ctx, const uint8_t *handles) {
	MC_REQUIRE(ctx && handles, MC_ERR_ARG, "bad arguments");
	McComm &cm = ctx->comm;
	MC_REQUIRE(cm.world, MC_ERR_STATE, "mc_comm_init has not been called");
	MC_CUDA(cudaSetDevice(ctx->device));
	for (int p = 0; p < cm.world; p++) {
		if (p == cm.rank) continue;
		cudaIpcMemHandle_t h;
		memcpy(&h, handles + (size_t)p * MC_COMM_HANDLE_BYTES, sizeof(h));
		void *ptr = nullptr;
		MC_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
		cm.peer_inbox[p] = (uint8_t *)ptr;
		cm.ipc_opened[p] = true;
	}
	cm.connected = true;
	return MC_OK;
}

extern "C" int mc_comm_connect_local(mc_ctx *const *ctxs, int world) {
	MC_REQUIRE(ctxs && world >= 1 && world <= MC_MAX_PEERS, MC_ERR_ARG, "bad arguments");
	for (int r = 0; r < world; r++) {
		MC_REQUIRE(ctxs[r] && ctxs[r]->comm.world == world && ctxs[r]->comm.rank == r, MC_ERR_STATE, "context %d: mc_comm_init(rank %d, world %d) first", r, r, world);
		// every rank leaves num_sms records per scan and expects as many from each peer
		MC_REQUIRE(ctxs[r]->num_sms == ctxs[0]->num_sms, MC_ERR_UNSUPPORTED, "GPUs %d and %d have different SM counts (%d, %d)", ctxs[0]->device, ctxs[r]->device, ctxs[0]->num_sms, ctxs[r]->num_sms);
	}
	for (int r = 0; r < world; r++) {
		MC_CUDA(cudaSetDevice(ctxs[r]->device));
		for (int p = 0; p < world; p++) {
			if (p == r) continue;
			if (ctxs[p]->device != ctxs[r]->device) {
				int can = 0;
				MC_CUDA(cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[p]->device));
				MC_REQUIRE(can, MC_ERR_UNSUPPORTED, "GPU %d cannot address GPU %d's memory (no NVLink / P2P)", ctxs[r]->device, ctxs[p]->device);
				cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[p]->device, 0);
				if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
				else MC_CUDA(e);
			}
			ctxs[r]->comm.peer_inbox[p] = ctxs[p]->comm.inbox;
		}
		ctxs[r]->comm.connected = true;
	}
	return MC_OK;
}


static int sharded_enqueue(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked, int slot) {
	MC_REQUIRE(ctx && ctx->have_hist, MC_ERR_STATE, "mc_scan_sharded_enqueue: histograms are not built");
	MC_REQUIRE(ctx->model.valid, MC_ERR_STATE, "mc_scan_sharded_enqueue: mc_set_model has not been called");
	McComm &cm = ctx->comm;
	MC_REQUIRE(cm.world && cm.connected, MC_ERR_STATE, "mc_scan_sharded_enqueue: mc_comm_init / mc_comm_connect first");
	MC_REQUIRE(slot >= 0 && slot < MC_XSLOTS, MC_ERR_ARG, "slot %d out of range (0..%d)", slot, MC_XSLOTS - 1);
	MC_REQUIRE(!cm.slot_pending[slot], MC_ERR_STATE, "slot %d still holds an exchange that was not collected", slot);
	MC_REQUIRE(center_row >= 0 && center_row < ctx->n, MC_ERR_ARG, "center row out of range");
	MC_REQUIRE(hi < lo || (lo >= 0 && hi < ctx->n), MC_ERR_ARG, "scan range [%lld,%lld] out of range", (long long)lo, (long long)hi);
	int rc = mc_comm_scan_push(ctx, center_row, lo, hi, remove_marked, slot, 0);
	if (rc) return rc;
	cm.slot_pending[slot] = 1;
	return MC_OK;
}

extern "C" int mc_scan_sharded_enqueue(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked, int slot) {
	return sharded_enqueue(ctx, center_row, lo, hi, remove_marked, slot);
}

extern "C" int mc_scan_sharded_enqueue_many(mc_ctx *ctx, const int64_t *center_rows, const int64_t *lo, const int64_t *hi,
                                            int count, int remove_marked, int slot0) {
	MC_REQUIRE(ctx && center_rows && lo && hi && count > 0, MC_ERR_ARG, "mc_scan_sharded_enqueue_many: bad arguments");
	for (int i = 0; i < count; i++) {
		const int rc = sharded_enqueue(ctx, center_rows[i], lo[i], hi[i], remove_marked, slot0 + i);
		if (rc) return rc;
	}
	return MC_OK;
}

static unsigned long long slot_offset(unsigned int epoch, int slot) {
	return ((unsigned long long)(epoch & 1u) * MC_XSLOTS + (unsigned long long)slot) * MC_MAX_PEERS * MC_SCAN_PARTS * MC_LL_RECORD_BYTES;
}

// internal: sharded scan of one slot on one rank; the kernel sends its own CTA partials when it ends.
// fence = 1 additionally orders earlier peer stores (marks written into another rank's array) before
// the record.
int mc_comm_scan_push(mc_ctx *ctx, int64_t center_row, int64_t lo, int64_t hi, int remove_marked, int slot, int fence) {
	McComm &cm = ctx->comm;
	MC_CUDA(cudaSetDevice(ctx->device));
	if (!ctx->d_scan_slots) {
		MC_CUDA(cudaMalloc(&ctx->d_scan_slots, (size_t)MC_SCAN_SLOTS * MC_SCAN_PARTS * sizeof(mc_scan_result)));
		for (int i = 0; i < MC_SCAN_SLOTS; i++) ctx->slot_nparts[i] = 0;
	}
	unsigned int epoch = ++cm.slot_epoch[slot];
	if (epoch == 0) epoch = cm.slot_epoch[slot] = 2;   // never 0 (the cleared inbox), parity kept
	McPeerPush push{};
	for (int p = 0; p < cm.world; p++) push.inbox[p] = (unsigned long long)cm.peer_inbox[p];
	push.world = cm.world;
	push.rank = cm.rank;
	push.epoch = epoch;
	push.fence = (unsigned int)fence;
	push.slot_off = slot_offset(epoch, slot);
	push.tiles_only = 0;
	return mc_launch_scan_push(ctx, center_row, lo, hi, remove_marked,
	                           (uint8_t *)ctx->d_scan_slots + (size_t)slot * MC_SCAN_PARTS * sizeof(mc_scan_result),
	                           &ctx->slot_nparts[slot], &push);
}

// internal: fold one slot on the device; *rec_dev_out is the combined record, *err_dev_out the timeout flag
int mc_comm_combine_dev(mc_ctx *ctx, int slot, const void **rec_dev_out, unsigned int **err_dev_out) {
	McComm &cm = ctx->comm;
	CombineArgs args;
	memset(&args, 0, sizeof(args));
	args.epoch[slot] = cm.slot_epoch[slot];
	args.slot_off[slot] = slot_offset(cm.slot_epoch[slot], slot);
	MC_CUDA(cudaSetDevice(ctx->device));
	mc_scan_result *d_out = (mc_scan_result *)cm.d_out;
	unsigned int *d_err = (unsigned int *)((uint8_t *)cm.d_out + (size_t)MC_XSLOTS * sizeof(mc_scan_result));
	scan_combine_kernel<<<1, COMBINE_THREADS, 0, ctx->stream>>>(cm.inbox, cm.world, ctx->num_sms, args, slot, d_out, d_err);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	*rec_dev_out = d_out + slot;
	*err_dev_out = d_err;
	return MC_OK;
}

// the device half of a collect: fold on the device, results to pinned host memory, an event behind it
extern "C" int mc_scan_sharded_combine(mc_ctx *ctx, int slot0, int nslots) {
	MC_REQUIRE(ctx, MC_ERR_ARG, "bad arguments");
	McComm &cm = ctx->comm;
	MC_REQUIRE(cm.world && cm.connected, MC_ERR_STATE, "mc_scan_sharded_combine: mc_comm_init / mc_comm_connect first");
	MC_REQUIRE(slot0 >= 0 && nslots > 0 && slot0 + nslots <= MC_XSLOTS, MC_ERR_ARG, "slot range invalid");
	CombineArgs args;
	memset(&args, 0, sizeof(args));
	for (int s = slot0; s < slot0 + nslots; s++) {
		MC_REQUIRE(cm.slot_pending[s] == 1, MC_ERR_STATE, "slot %d has no exchange in flight", s);
		args.epoch[s] = cm.slot_epoch[s];
		args.slot_off[s] = slot_offset(cm.slot_epoch[s], s);
	}
	MC_CUDA(cudaSetDevice(ctx->device));
	if (!cm.h_out) {
		MC_CUDA(cudaMallocHost(&cm.h_out, (size_t)MC_XSLOTS * sizeof(mc_scan_result) + 64));
		MC_CUDA(cudaEventCreateWithFlags(&cm.done, cudaEventDisableTiming));
	}
	mc_scan_result *d_out = (mc_scan_result *)cm.d_out;
	unsigned int *d_err = (unsigned int *)((uint8_t *)cm.d_out + (size_t)MC_XSLOTS * sizeof(mc_scan_result));
	scan_combine_kernel<<<nslots, COMBINE_THREADS, 0, ctx->stream>>>(cm.inbox, cm.world, ctx->num_sms, args, slot0, d_out, d_err);
	ctx->launches++;
	MC_CUDA(cudaGetLastError());
	mc_scan_result *h_out = (mc_scan_result *)cm.h_out;
	MC_CUDA(cudaMemcpyAsync(h_out + slot0, d_out + slot0, (size_t)nslots * sizeof(mc_scan_result), cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaMemcpyAsync(h_out + MC_XSLOTS, d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
	MC_CUDA(cudaEventRecord(cm.done, ctx->stream));
	for (int s = slot0; s < slot0 + nslots; s++) cm.slot_pending[s] = 2;
	return MC_OK;
}

// the host half: wait for the event of the last combine and hand the summaries out
extern "C" int mc_scan_sharded_wait(mc_ctx *ctx, int slot0, int nslots, mc_scan_result *res) {
	MC_REQUIRE(ctx && res, MC_ERR_ARG, "bad arguments");
	McComm &cm = ctx->comm;
	MC_REQUIRE(slot0 >= 0 && nslots > 0 && slot0 + nslots <= MC_XSLOTS, MC_ERR_ARG, "slot range invalid");
	for (int s = slot0; s < slot0 + nslots; s++) MC_REQUIRE(cm.slot_pending[s] == 2, MC_ERR_STATE, "slot %d was not combined", s);
	MC_CUDA(cudaSetDevice(ctx->device));
	MC_CUDA(cudaEventSynchronize(cm.done));
	for (int s = slot0; s < slot0 + nslots; s++) cm.slot_pending[s] = 0;
	const mc_scan_result *h_out = (const mc_scan_result *)cm.h_out;
	if (*(const unsigned int *)(h_out + MC_XSLOTS)) {
		unsigned int *d_err = (unsigned int *)((uint8_t *)cm.d_out + (size_t)MC_XSLOTS * sizeof(mc_scan_result));
		MC_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned int), ctx->stream));
		mc_set_error("sharded scan: a peer's records did not arrive within %.0f s (rank %d of %d)", COMBINE_TIMEOUT_NS * 1e-9, cm.rank, cm.world);
		return MC_ERR_CUDA;
	}
	memcpy(res, h_out + slot0, (size_t)nslots * sizeof(mc_scan_result));
	return MC_OK;
}

extern "C" int mc_scan_sharded_collect(mc_ctx *ctx, int slot0, int nslots, mc_scan_result *res) {
	const int rc = mc_scan_sharded_combine(ctx, slot0, nslots);
	if (rc) return rc;
	return mc_scan_sharded_wait(ctx, slot0, nslots, res);
}

// ---------------------------------------------------------------------------------------------
// Burst path (streaming callers, bench.py): the scans run back to back on the context's stream
// exactly like single-GPU scans (tiles-only variant, no peer store in the kernel); a second stream
// folds each scan's CTA partials to ONE record per rank, stores it into every inbox and combines the
// world records per scan.  The scan stream never waits for the exchange.
// ---------------------------------------------------------------------------------------------
struct BurstArgs {
	unsigned long long slot_off[MC_XSLOTS];
	unsigned int epoch[MC_XSLOTS];
	unsigned int tag[MC_XSLOTS];      // tag the CTA partials of this use of the slot carry
};

// one warp per slot: fold the CTA partials of this rank's scan, then send the record (CTA index 0)
__global__ void __launch_bounds__(32) fold_send_kernel(const unsigned int *__restrict__ ll_partials, int nparts, BurstArgs args, int slot0,
                                                       McPeerPush push, unsigned int *__restrict__ err) {
	const int slot = slot0 + blockIdx.x, lane = threadIdx.x;
	// this kernel is not ordered behind the scans by the stream: it polls the {data, tag} copies of the
	// CTA partials until every word carries this use's tag
	const unsigned int tag = args.tag[slot];
	const unsigned int *base = ll_partials + (size_t)slot * MC_SCAN_PARTS * 16;
	mc_scan_result b;
	b.n_eval = 0; b.n_pos = 0; b.best_row = -1; b.best_f0 = -1.0;
	unsigned long long t0;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
	bool failed = false;
	for (int i = lane; i < nparts && !failed; i += 32) {
		const unsigned int *rec = base + (size_t)i * 16;
		uint4 q[4];
		for (;;) {
#pragma unroll
			for (int j = 0; j < 4; j++) q[j] = ld_volatile16(rec + j * 4);
			bool ok = true;
#pragma unroll
			for (int j = 0; j < 4; j++) ok = ok && q[j].y == tag && q[j].w == tag;
			if (ok) break;
			__nanosleep(100);
			unsigned long long t1;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
			if (t1 - t0 > COMBINE_TIMEOUT_NS) { failed = true; break; }
		}
		if (failed) break;
		mc_scan_result r;
		r.n_eval = (long long)((unsigned long long)q[0].x | ((unsigned long long)q[0].z << 32));
		r.n_pos = (long long)((unsigned long long)q[1].x | ((unsigned long long)q[1].z << 32));
		r.best_row = (long long)((unsigned long long)q[2].x | ((unsigned long long)q[2].z << 32));
		r.best_f0 = __longlong_as_double((long long)((unsigned long long)q[3].x | ((unsigned long long)q[3].z << 32)));
		mc_scan_merge(b, r);
	}
	if (__any_sync(MC_FULL_MASK, failed)) { if (lane == 0) atomicExch(err, 1u); return; }
	mc_scan_warp_fold(b);
	unsigned long long f[4] = {(unsigned long long)b.n_eval, (unsigned long long)b.n_pos, (unsigned long long)b.best_row,
	                           (unsigned long long)__double_as_longlong(b.best_f0)};
	const int w = lane & 7;
	const unsigned long long fld = (w >> 1) == 0 ? f[0] : ((w >> 1) == 1 ? f[1] : ((w >> 1) == 2 ? f[2] : f[3]));
	const unsigned int data = (w & 1) ? (unsigned int)(fld >> 32) : (unsigned int)fld;
	for (int q = lane >> 3; q < push.world; q += 4) {
		const unsigned long long dst = push.inbox[q] + args.slot_off[slot] +
			((unsigned long long)push.rank * MC_SCAN_PARTS + 0) * MC_LL_RECORD_BYTES + (unsigned long long)w * 8;
		asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(data), "r"(args.epoch[slot]) : "memory");
	}
}

static int burst_streams(mc_ctx *ctx) {
	McComm &cm = ctx->comm;
	if (cm.xstream) return MC_OK;
	MC_CUDA(cudaMalloc(&cm.d_ll_partials, (size_t)MC_XSLOTS * MC_SCAN_PARTS * 16 * sizeof(unsigned int)));
	MC_CUDA(cudaMemset(cm.d_ll_partials, 0, (size_t)MC_XSLOTS * MC_SCAN_PARTS * 16 * sizeof(unsigned int)));
	// highest priority: when an SM has room, the (tiny) exchange kernels go before the queued CTAs of
	// the next scans -- otherwise the summaries of a burst would only leave once the NEXT burst has
	// drained (measured: 70 us from the end of a burst to its summaries on the host)
	int prio_lo = 0, prio_hi = 0;
	MC_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
	MC_CUDA(cudaStreamCreateWithPriority(&cm.xstream, cudaStreamNonBlocking, prio_hi));
	if (!cm.h_out) {
		MC_CUDA(cudaMallocHost(&cm.h_out, (size_t)MC_XSLOTS * sizeof(mc_scan_result) + 64));
		MC_CUDA(cudaEventCreateWithFlags(&cm.done, cudaEventDisableTiming));
	}
	return MC_OK;
}

// One pipelined burst: enqueue this burst's scans on the scan stream and their fold + send + combine
// on the exchange stream, then wait for the summaries of an EARLIER burst (slots prev_slot0..) -- the
// GPU already runs this burst while the host waits.  A burst's slots should start at a multiple of
// MC_SCAN_BATCH (one completion event per such bank, four banks).
extern "C" int mc_scan_sharded_burst(mc_ctx *ctx, const int64_t *center_rows, const int64_t *lo, const int64_t *hi,
                                     int count, int remove_marked, int slot0, int prev_slot0, int prev_count,
                                     mc_scan_result *prev_res) {
	MC_REQUIRE(ctx && ctx->have_hist && ctx->model.valid, MC_ERR_STATE, "mc_scan_sharded_burst: histograms / model missing");
	McComm &cm = ctx->comm;
	MC_REQUIRE(cm.world && cm.connected, MC_ERR_STATE, "mc_scan_sharded_burst: mc_comm_init / mc_comm_connect first");
	MC_REQUIRE(count >= 0 && prev_count >= 0 && slot0 >= 0 && slot0 + count <= MC_XSLOTS && prev_slot0 >= 0 && prev_slot0 + prev_count <= MC_XSLOTS,
	           MC_ERR_ARG, "slot ranges invalid");
	MC_REQUIRE(count == 0 || (center_rows && lo && hi), MC_ERR_ARG, "mc_scan_sharded_burst: bad arguments");
	MC_REQUIRE(prev_count == 0 || prev_res, MC_ERR_ARG, "mc_scan_sharded_burst: prev_res is NULL");
	MC_CUDA(cudaSetDevice(ctx->device));
	int rc = burst_streams(ctx);
	if (rc) return rc;
	static const bool dbg = getenv("MC_DEBUG_TIMING") != nullptr;
	static const bool no_exchange = getenv("MC_BURST_NO_EXCHANGE") != nullptr;   // diagnosis only: scans without fold / send / combine
	static double t_comb = 0, t_scan = 0, t_fold = 0, t_wait = 0;
	static long n_calls = 0;
	auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
	double tq0 = dbg ? now() : 0, tq1 = 0, tq2 = 0, tq3 = 0;
	mc_scan_result *d_out = (mc_scan_result *)cm.d_out;
	unsigned int *d_err = (unsigned int *)((uint8_t *)cm.d_out + (size_t)MC_XSLOTS * sizeof(mc_scan_result));
	mc_scan_result *h_out = (mc_scan_result *)cm.h_out;
	for (int sl = prev_slot0; sl < prev_slot0 + prev_count; sl++)
		MC_REQUIRE(cm.slot_pending[sl] == 4 || (no_exchange && cm.slot_pending[sl] == 3), MC_ERR_STATE, "slot %d holds no burst scan", sl);
	if (dbg) tq1 = now();
	if (count > 0) {
		if (!ctx->d_scan_slots) {
			MC_CUDA(cudaMalloc(&ctx->d_scan_slots, (size_t)MC_SCAN_SLOTS * MC_SCAN_PARTS * sizeof(mc_scan_result)));
			for (int i = 0; i < MC_SCAN_SLOTS; i++) ctx->slot_nparts[i] = 0;
		}
		McPeerPush push{};
		for (int p = 0; p < cm.world; p++) push.inbox[p] = (unsigned long long)cm.peer_inbox[p];
		push.world = cm.world;
		push.rank = cm.rank;
		push.tiles_only = 1;
		BurstArgs bargs;
		memset(&bargs, 0, sizeof(bargs));
		for (int i = 0; i < count; i++) {
			const int slot = slot0 + i;
			MC_REQUIRE(!cm.slot_pending[slot], MC_ERR_STATE, "slot %d still holds an exchange that was not collected", slot);
			MC_REQUIRE(center_rows[i] >= 0 && center_rows[i] < ctx->n, MC_ERR_ARG, "center row out of range");
			MC_REQUIRE(hi[i] < lo[i] || (lo[i] >= 0 && hi[i] < ctx->n), MC_ERR_ARG, "scan range out of range");
		}
		// scans that remove nothing are independent: up to MC_SCAN_BATCH per launch; otherwise one
		// launch per scan, chained by programmatic dependent launch
		const int per_launch = remove_marked ? 1 : MC_SCAN_BATCH;
		for (int i0 = 0; i0 < count; i0 += per_launch) {
			const int m = count - i0 < per_launch ? count - i0 : per_launch;
			McScanReq req[MC_SCAN_BATCH];
			for (int i = 0; i < m; i++) {
				const int slot = slot0 + i0 + i;
				unsigned int epoch = ++cm.slot_epoch[slot];
				if (epoch == 0) epoch = cm.slot_epoch[slot] = 2;
				bargs.epoch[slot] = epoch;
				bargs.slot_off[slot] = slot_offset(epoch, slot);
				req[i].lo = lo[i0 + i]; req[i].hi = hi[i0 + i]; req[i].center_row = center_rows[i0 + i];
				req[i].partials_dev = (uint8_t *)ctx->d_scan_slots + (size_t)slot * MC_SCAN_PARTS * sizeof(mc_scan_result);
				req[i].marks_dev = nullptr;
				cm.slot_uses[slot]++;
				if (cm.slot_uses[slot] == 0) cm.slot_uses[slot] = 1;   // never the tag of the cleared buffer
				req[i].ll_partials_dev = cm.d_ll_partials + (size_t)slot * MC_SCAN_PARTS * 16;
				req[i].ll_tag = cm.slot_uses[slot];
				bargs.tag[slot] = cm.slot_uses[slot];
				cm.slot_pending[slot] = 3;
			}
			rc = mc_launch_scan_batch(ctx, req, m, remove_marked & MC_SCAN_REMOVE, &ctx->slot_nparts[slot0 + i0], &push);
			if (rc) return rc;
		}
		if (dbg) tq2 = now();
		if (!no_exchange) {
			fold_send_kernel<<<count, 32, 0, cm.xstream>>>(cm.d_ll_partials, ctx->num_sms, bargs, slot0, push, d_err);
			ctx->launches++;
			MC_CUDA(cudaGetLastError());
			// the combine of THIS burst goes out right behind its send (one warp per scan: world records
			// to fold, and a CTA this small fits next to the resident scan CTAs); its summaries are on the
			// host a few microseconds after the burst's last scan, whenever the caller asks for them
			CombineArgs args;
			memset(&args, 0, sizeof(args));
			for (int sl = slot0; sl < slot0 + count; sl++) {
				args.epoch[sl] = cm.slot_epoch[sl];
				args.slot_off[sl] = slot_offset(cm.slot_epoch[sl], sl);
				cm.slot_pending[sl] = 4;
			}
			scan_combine_kernel<<<count, 32, 0, cm.xstream>>>(cm.inbox, cm.world, 1, args, slot0, d_out, d_err);
			ctx->launches++;
			MC_CUDA(cudaGetLastError());
			MC_CUDA(cudaMemcpyAsync(h_out + slot0, d_out + slot0, (size_t)count * sizeof(mc_scan_result), cudaMemcpyDeviceToHost, cm.xstream));
			MC_CUDA(cudaMemcpyAsync(h_out + MC_XSLOTS, d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, cm.xstream));
			cudaEvent_t &ev = cm.burst_done[(slot0 / MC_SCAN_BATCH) & 3];
			if (!ev) MC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
			MC_CUDA(cudaEventRecord(ev, cm.xstream));
		}
	}
	if (dbg) { if (tq2 == 0) tq2 = now(); tq3 = now(); }
	if (prev_count > 0 && no_exchange) {
		for (int s = prev_slot0; s < prev_slot0 + prev_count; s++) cm.slot_pending[s] = 0;
		memset(prev_res, 0, (size_t)prev_count * sizeof(mc_scan_result));
		return MC_OK;
	}
	if (prev_count > 0) {
		cudaEvent_t ev = cm.burst_done[(prev_slot0 / MC_SCAN_BATCH) & 3];
		MC_REQUIRE(ev, MC_ERR_STATE, "no burst was enqueued on slots %d..", prev_slot0);
		MC_CUDA(cudaEventSynchronize(ev));
		if (dbg) {
			t_comb += tq1 - tq0; t_scan += tq2 - tq1; t_fold += tq3 - tq2; t_wait += now() - tq3;
			if (++n_calls % 25 == 0)
				fprintf(stderr, "[mc_scan_sharded_burst rank %d] %ld calls: combine enqueue %.1f us, scans enqueue %.1f us, fold enqueue %.1f us, wait for the previous burst %.1f us (averages)\n",
				        cm.rank, n_calls, t_comb / n_calls * 1e6, t_scan / n_calls * 1e6, t_fold / n_calls * 1e6, t_wait / n_calls * 1e6);
		}
		for (int s = prev_slot0; s < prev_slot0 + prev_count; s++) cm.slot_pending[s] = 0;
		if (*(const unsigned int *)(h_out + MC_XSLOTS)) {
			MC_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned int), cm.xstream));
			mc_set_error("sharded burst: a peer's records did not arrive within %.0f s (rank %d of %d)", COMBINE_TIMEOUT_NS * 1e-9, cm.rank, cm.world);
			return MC_ERR_CUDA;
		}
		memcpy(prev_res, h_out + prev_slot0, (size_t)prev_count * sizeof(mc_scan_result));
	}
	return MC_OK;
}
