// comm.cu — multi-GPU plumbing for a row-partitioned system: one process per GPU, every rank
// holds a block of consecutive rows of A and the matching slices of all vectors.  Two exchange steps exist on the
// path, both stream-ordered so that the iteration stays asynchronous to the host:
//   halo       before every SpMV the entries of the input vector that other ranks' rows reference are sent to
//              them (ncclSend/ncclRecv in one group over NVLink) and land in the ghost tail of the extended vector
//              [n_local local entries | ghosts of peer 0 | ghosts of peer 1 | ...];
//   allreduce  the per-rank totals a fused kernel left in DevState::red are summed in place (ncclAllReduce on
//              1-8 doubles), after which k_finish runs the scalar epilogue identically on every rank.
// The reference has no distributed path at all ("Parallelism strategies: none").
//
// NCCL is resolved with dlopen at first use (the copy already loaded by the host process, e.g. torch's, wins), so
// single-GPU users of liblcgb200.so do not need it.
#include "engine.cuh"
#include "../../include/lcgb200.h"
#include <dlfcn.h>
#include <cstring>
#include <vector>
#include <string>
#include <algorithm>

namespace lcgb200 {

namespace {

struct NcclId { char internal[128]; };
typedef void* NcclCommT;

struct NcclApi {
	void* lib = nullptr;
	int (*GetUniqueId)(NcclId*) = nullptr;
	int (*CommInitRank)(NcclCommT*, int, NcclId, int) = nullptr;
	int (*CommDestroy)(NcclCommT) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, NcclCommT, cudaStream_t) = nullptr;
	int (*Send)(const void*, size_t, int, int, NcclCommT, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, NcclCommT, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;

	bool load()
	{
		if (AllReduce) return true;
		const char* names[] = {"libnccl.so.2", "libnccl.so"};
		for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (lib) break; }
		if (!lib) for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
		// torch bundles NCCL under a path that is not on the loader's search list but is already mapped: look the
		// symbols up in the global scope as a last resort
		void* scope = lib ? lib : RTLD_DEFAULT;
#define LCG_NCCL_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(scope, name))
		LCG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
		LCG_NCCL_SYM(CommInitRank, "ncclCommInitRank");
		LCG_NCCL_SYM(CommDestroy, "ncclCommDestroy");
		LCG_NCCL_SYM(AllReduce, "ncclAllReduce");
		LCG_NCCL_SYM(Send, "ncclSend");
		LCG_NCCL_SYM(Recv, "ncclRecv");
		LCG_NCCL_SYM(GroupStart, "ncclGroupStart");
		LCG_NCCL_SYM(GroupEnd, "ncclGroupEnd");
		LCG_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef LCG_NCCL_SYM
		if (!(GetUniqueId && CommInitRank && CommDestroy && AllReduce && Send && Recv && GroupStart && GroupEnd))
		{
			AllReduce = nullptr;
			set_error_msg("NCCL (libnccl.so.2) could not be loaded: multi-GPU solves need it");
			return false;
		}
		return true;
	}
};
NcclApi g_nccl;

void nccl_check(int rc, const char* what)
{
	if (rc == 0) return;
	std::string msg = std::string(what) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
	set_error_msg(msg.c_str());
	throw CudaFailure();
}

// NVLink transport, one kernel per halo exchange (grid <= SM count, so all blocks are co-resident):
//   1. push my boundary entries straight into the peers' mailboxes (remote stores over NVLink);
//   2. every block fences at system scope and takes a ticket; the last one bumps my sequence flag in each peer's window;
//   3. every block waits until the peers' flags for this exchange have arrived in MY window and copies its share of my
//      mailbox into the ghost tail of x — the SpMV that follows is the unmodified single-GPU kernel.
// Pushing never waits on anything, so two ranks running this kernel cannot block each other.
template <class T>
__global__ void __launch_bounds__(256) k_halo_exchange(CommDev* c, T* __restrict__ x, DevState* st)
{
	pdl_enter();
	if (st_done(st)) return;
	const unsigned long long seq = c->halo_seq + 1;
	const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
	for (int p = 0; p < c->n_peers; p++)
	{
		const int cnt = c->send_count[p];
		if (cnt <= 0) continue;
		T* dst = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(c->win[c->peer_rank[p]]) + kMailboxOffset)
			+ (size_t)(seq & 1) * (size_t)c->remote_ghost[p] + (size_t)c->remote_off[p];
		const int first = c->send_first[p], off = c->send_off[p];
		const bool contig = c->contiguous[p] != 0;
		for (int i = gtid; i < cnt; i += gsz) dst[i] = contig ? x[first + i] : x[c->send_idx[off + i]];
	}
	__threadfence_system();
	__shared__ int s_flag;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		const bool last = atomicAdd(&c->ticket, 1u) == gridDim.x - 1;
		if (last)
		{
			__threadfence_system();
			for (int p = 0; p < c->n_peers; p++)
				if (c->send_count[p] > 0) st_relaxed_sys(&c->win[c->peer_rank[p]]->halo_flag[c->rank], seq);
		}
		bool ok = true;
		for (int p = 0; p < c->n_peers && ok; p++)
			if (c->recv_count[p] > 0) ok = spin_until(&c->win[c->rank]->halo_flag[c->peer_rank[p]], seq, st->spin_timeout_ns);
		if (!ok) { st->ret = RC_UNKNOWN; st->done = 1; c->abort_flag = 1; }
		s_flag = ok ? 1 : 0;
	}
	__syncthreads();
	if (s_flag)
	{
		const T* mail = reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(c->win[c->rank]) + kMailboxOffset) + (size_t)(seq & 1) * (size_t)c->n_ghost;
		T* ghost = x + c->n_local;
		for (int i = gtid; i < c->n_ghost; i += gsz) ghost[i] = __ldcv(mail + i);
	}
	// the block that leaves last publishes the new sequence number for the next exchange
	__syncthreads();
	if (threadIdx.x == 0 && atomicAdd(&c->ticket2, 1u) == gridDim.x - 1) { c->halo_seq = seq; c->ticket = 0u; c->ticket2 = 0u; }
}

// NVLink transport, push half only (the receive half lives in k_spmv: boundary tiles wait for the flags and read the
// mailbox in place).  Used when the SpMV input was not produced by one of our pushing kernels: the first SpMV of a solve
// (the caller's m), stand-alone lcgb200_csr_spmv, send lists that are not runs of consecutive rows.  Before overwriting
// mailbox buffer (seq & 1) the sender waits until every receiver has acknowledged exchange seq - 2 — inside the solvers
// that is already implied by the reductions between two exchanges; back-to-back stand-alone SpMVs need it.
template <class T>
__global__ void __launch_bounds__(256) k_halo_push(CommDev* c, const T* __restrict__ x, DevState* st)
{
	pdl_enter();
	if (st_done(st)) return;
	const unsigned long long seq = c->halo_seq + 1;
	__shared__ int s_ok;
	if (threadIdx.x == 0)
	{
		bool ok = true;
		if (seq > 2)
			for (int p = 0; p < c->n_peers && ok; p++)
				if (c->send_count[p] > 0) ok = spin_until(&c->win[c->rank]->halo_ack[c->peer_rank[p]], seq - 2, st->spin_timeout_ns);
		if (!ok) { st->ret = RC_UNKNOWN; st->done = 1; c->abort_flag = 1; }
		s_ok = ok ? 1 : 0;
	}
	__syncthreads();
	bool pushed = false;
	if (s_ok)
	{
		const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
		for (int p = 0; p < c->n_peers; p++)
		{
			const int cnt = c->send_count[p];
			if (cnt <= 0) continue;
			T* dst = mailbox_of<T>(c->win[c->peer_rank[p]], seq, c->remote_ghost[p]) + c->remote_off[p];
			const int first = c->send_first[p], off = c->send_off[p];
			const bool contig = c->contiguous[p] != 0;
			for (int i = gtid; i < cnt; i += gsz) { dst[i] = contig ? x[first + i] : x[c->send_idx[off + i]]; pushed = true; }
		}
	}
	push_signal(c, seq, pushed);
}

// flag[tile] = 1 if any entry of the tile's rows references a ghost column (>= n_local); one warp per tile
__global__ void k_tile_boundary(int n_tiles, const int4* __restrict__ tiles, const int* __restrict__ row_ptr, const int* __restrict__ col, int n_local, int* flag)
{
	const int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (tile >= n_tiles) return;
	const int4 td = tiles[tile];
	bool hit = false;
	for (int k = row_ptr[td.x] + lane; k < td.w && !hit; k += 32) hit = col[k] >= n_local;
	hit = __any_sync(0xffffffffu, hit);
	if (lane == 0) flag[tile] = hit ? 1 : 0;
}

// send_buf[i] = x[idx[i]] for the entries that are not sent in place
template <class T>
__global__ void k_halo_pack(const T* __restrict__ x, const int* __restrict__ idx, T* __restrict__ out, int count)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) out[i] = x[idx[i]];
}

}  // namespace

class NcclComm : public Comm {
public:
	NcclCommT comm = nullptr;
	int rank_ = 0, size_ = 1;
	struct Peer { int rank; int send_count; int send_off; bool contiguous; int send_first; int recv_count; int recv_off; };
	std::vector<Peer> peers;
	int n_local = 0, n_ghost = 0, n_packed = 0;
	int* d_send_idx = nullptr;     // concatenated local indices of the packed peers
	void* d_send_buf = nullptr;    // packed values (16 bytes per entry: enough for double2)
	int halos = 0, allreduces = 0;
	// NVLink peer-memory transport (optional; needs CUDA IPC between the ranks' devices)
	CommWindow* window = nullptr; size_t window_bytes = 0;
	CommDev* d_dev = nullptr; CommDev h_dev;
	void* peer_map[kMaxRanks] = {nullptr};
	bool p2p_ready = false;
	int push_grid = 1;

	CommDev* dev() override { return p2p_ready ? d_dev : nullptr; }
	bool fused_push_ok() const override { return p2p_ready && h_dev.all_contiguous != 0; }
	void push(const void* x, int elem_bytes, cudaStream_t s, DevState* st) override
	{
		if (peers.empty()) return;
		if (elem_bytes == 8) launch_k(k_halo_push<double>, push_grid, 256, 0, s, d_dev, (const double*)x, st);
		else launch_k(k_halo_push<double2>, push_grid, 256, 0, s, d_dev, (const double2*)x, st);
		halos++;
	}
	// after a solve on the NVLink transport: did a cross-GPU wait time out?  A timed-out communicator is poisoned — its
	// sequence counters and mailboxes can no longer be trusted to agree with the peers'.
	bool poisoned_ = false;
	bool poisoned() const override { return poisoned_; }
	bool check_abort() override
	{
		if (!p2p_ready || !d_dev) return false;
		unsigned int flag = 0;
		if (cudaMemcpy(&flag, reinterpret_cast<const char*>(d_dev) + offsetof(CommDev, abort_flag), sizeof(flag), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
		if (flag) poisoned_ = true;
		return flag != 0;
	}

	~NcclComm() override
	{
		for (int r = 0; r < kMaxRanks; r++) if (peer_map[r]) cudaIpcCloseMemHandle(peer_map[r]);
		cudaFree(window); cudaFree(d_dev);
		cudaFree(d_send_idx); cudaFree(d_send_buf);
		if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
	}
	int rank() const override { return rank_; }
	int size() const override { return size_; }

	void allreduce(double* dev, int count, cudaStream_t s) override
	{
		nccl_check(g_nccl.AllReduce(dev, dev, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, comm, s), "ncclAllReduce");
		allreduces++;
	}

	void halo(void* x_ext, int elem_bytes, cudaStream_t s, bool p2p, DevState* st) override
	{
		if (peers.empty()) return;
		if (p2p && p2p_ready)
		{
			if (elem_bytes == 8) launch_k(k_halo_exchange<double>, push_grid, 256, 0, s, d_dev, (double*)x_ext, st);
			else launch_k(k_halo_exchange<double2>, push_grid, 256, 0, s, d_dev, (double2*)x_ext, st);
			halos++;
			return;
		}
		char* x = static_cast<char*>(x_ext);
		if (n_packed > 0)
		{
			const int grid = (n_packed + 255) / 256;
			if (elem_bytes == 8) k_halo_pack<double><<<grid, 256, 0, s>>>((const double*)x_ext, d_send_idx, (double*)d_send_buf, n_packed);
			else k_halo_pack<double2><<<grid, 256, 0, s>>>((const double2*)x_ext, d_send_idx, (double2*)d_send_buf, n_packed);
		}
		nccl_check(g_nccl.GroupStart(), "ncclGroupStart");
		for (const Peer& p : peers)
		{
			if (p.send_count > 0)
			{
				const void* src = p.contiguous ? (const void*)(x + (size_t)p.send_first * elem_bytes)
				                               : (const void*)((char*)d_send_buf + (size_t)p.send_off * elem_bytes);
				nccl_check(g_nccl.Send(src, (size_t)p.send_count * elem_bytes, /*ncclChar*/ 0, p.rank, comm, s), "ncclSend");
			}
			if (p.recv_count > 0)
				nccl_check(g_nccl.Recv(x + ((size_t)n_local + p.recv_off) * elem_bytes, (size_t)p.recv_count * elem_bytes, 0, p.rank, comm, s), "ncclRecv");
		}
		nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
		halos++;
	}
};

}  // namespace lcgb200

using namespace lcgb200;

namespace {
template <class F> int guarded_comm(F&& f)
{
	try { return f(); }
	catch (const CudaFailure&) { return LCGB200_UNKNOWN_ERROR; }
	catch (const std::exception& e) { set_error_msg(e.what()); return LCGB200_UNKNOWN_ERROR; }
}
}  // namespace

extern "C" {

int lcgb200_comm_unique_id(void* id_out, int capacity)
{
	if (!id_out || capacity < LCGB200_COMM_ID_BYTES) return LCGB200_INVALID_POINTER;
	if (!g_nccl.load()) return LCGB200_UNKNOWN_ERROR;
	NcclId id;
	int rc = g_nccl.GetUniqueId(&id);
	if (rc != 0) { set_error_msg("ncclGetUniqueId failed"); return LCGB200_UNKNOWN_ERROR; }
	std::memcpy(id_out, &id, sizeof(id));
	return 0;
}

int lcgb200_comm_create(lcgb200_comm_t* out, int rank, int size, const void* unique_id)
{
	if (!out || !unique_id) return LCGB200_INVALID_POINTER;
	*out = nullptr;
	if (size < 1 || rank < 0 || rank >= size) return LCGB200_INVILAD_VARIABLE_SIZE;
	if (!g_nccl.load()) return LCGB200_UNKNOWN_ERROR;
	NcclComm* c = new NcclComm();
	int rc = guarded_comm([&]() {
		NcclId id; std::memcpy(&id, unique_id, sizeof(id));
		c->rank_ = rank; c->size_ = size;
		nccl_check(g_nccl.CommInitRank(&c->comm, size, id, rank), "ncclCommInitRank");
		return 0;
	});
	if (rc != 0) { delete c; return rc; }
	*out = reinterpret_cast<lcgb200_comm_t>(c);
	return 0;
}

int lcgb200_comm_destroy(lcgb200_comm_t comm)
{
	delete reinterpret_cast<NcclComm*>(comm);
	return 0;
}

int lcgb200_comm_stats(lcgb200_comm_t comm, int* halos, int* allreduces)
{
	NcclComm* c = reinterpret_cast<NcclComm*>(comm);
	if (!c) return LCGB200_INVALID_POINTER;
	if (halos) *halos = c->halos;
	if (allreduces) *allreduces = c->allreduces;
	return 0;
}

int lcgb200_csr_set_partition(lcgb200_csr_t A, lcgb200_comm_t comm, long long n_global, int n_peers, const int* peer_ranks,
	const int* send_counts, const int* send_idx, const int* recv_counts)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	NcclComm* c = reinterpret_cast<NcclComm*>(comm);
	if (!h || !c) return LCGB200_INVALID_POINTER;
	if (n_peers < 0 || n_global < h->n_rows) return LCGB200_INVILAD_VARIABLE_SIZE;
	if (n_peers > 0 && (!peer_ranks || !send_counts || !recv_counts)) return LCGB200_INVALID_POINTER;
	if (h->t_row_ptr) { set_error_msg("the local transpose of a rectangular row block is not A^T: build the rows of A^T as a second partitioned handle and attach it (lcgb200_csr_attach_transpose)"); return LCGB200_SIZE_NOT_MATCH; }
	return guarded_comm([&]() {
		c->peers.clear();
		c->n_local = h->n_rows;
		std::vector<int> packed;
		int recv_off = 0, send_pos = 0;
		for (int p = 0; p < n_peers; p++)
		{
			NcclComm::Peer pe;
			pe.rank = peer_ranks[p]; pe.send_count = send_counts[p]; pe.recv_count = recv_counts[p];
			pe.recv_off = recv_off; recv_off += recv_counts[p];
			pe.contiguous = true; pe.send_first = pe.send_count > 0 ? send_idx[send_pos] : 0; pe.send_off = 0;
			for (int i = 0; i < pe.send_count; i++)
			{
				const int v = send_idx[send_pos + i];
				if (v < 0 || v >= h->n_rows) { set_error_msg("send index outside the local rows"); throw CudaFailure(); }
				if (v != pe.send_first + i) pe.contiguous = false;
			}
			if (!pe.contiguous)
			{
				pe.send_off = (int)packed.size();
				packed.insert(packed.end(), send_idx + send_pos, send_idx + send_pos + pe.send_count);
			}
			send_pos += pe.send_count;
			if (pe.rank < 0 || pe.rank >= c->size_ || pe.rank == c->rank_) { set_error_msg("bad peer rank"); throw CudaFailure(); }
			c->peers.push_back(pe);
		}
		if (h->n_rows + recv_off != h->n_cols) { set_error_msg("n_rows + ghost entries must equal n_cols of the rectangular block"); throw CudaFailure(); }
		c->n_ghost = recv_off;
		c->n_packed = (int)packed.size();
		cudaFree(c->d_send_idx); cudaFree(c->d_send_buf); c->d_send_idx = nullptr; c->d_send_buf = nullptr;
		if (c->n_packed > 0)
		{
			LCG_CUDA_CHECK(cudaMalloc((void**)&c->d_send_idx, packed.size() * sizeof(int)));
			LCG_CUDA_CHECK(cudaMalloc(&c->d_send_buf, packed.size() * 16));
			LCG_CUDA_CHECK(cudaMemcpy(c->d_send_idx, packed.data(), packed.size() * sizeof(int), cudaMemcpyHostToDevice));
		}
		h->comm = c;
		h->n_global = n_global;
		// window for the NVLink transport: header + double-buffered mailbox of 16-byte slots
		c->p2p_ready = false;
		for (int r = 0; r < kMaxRanks; r++) if (c->peer_map[r]) { cudaIpcCloseMemHandle(c->peer_map[r]); c->peer_map[r] = nullptr; }
		cudaFree(c->window); c->window = nullptr; cudaFree(c->d_dev); c->d_dev = nullptr;
		if (c->size_ <= kMaxRanks && n_peers <= kMaxRanks)
		{
			c->window_bytes = kMailboxOffset + 2 * (size_t)std::max(c->n_ghost, 1) * 16;
			LCG_CUDA_CHECK(cudaMalloc((void**)&c->window, c->window_bytes));
			LCG_CUDA_CHECK(cudaMemset(c->window, 0, c->window_bytes));
			LCG_CUDA_CHECK(cudaMalloc((void**)&c->d_dev, sizeof(CommDev)));
			CommDev& d = c->h_dev;
			std::memset(&d, 0, sizeof(d));
			d.rank = c->rank_; d.size = c->size_; d.n_local = c->n_local; d.n_ghost = c->n_ghost; d.n_peers = n_peers; d.n_packed = c->n_packed;
			d.send_idx = c->d_send_idx;
			int max_send = 1;
			for (int p = 0; p < n_peers; p++)
			{
				const NcclComm::Peer& pe = c->peers[(size_t)p];
				d.peer_rank[p] = pe.rank; d.send_first[p] = pe.send_first; d.send_count[p] = pe.send_count; d.send_off[p] = pe.send_off;
				d.contiguous[p] = pe.contiguous ? 1 : 0; d.recv_count[p] = pe.recv_count;
				max_send = std::max(max_send, pe.send_count);
			}
			c->push_grid = std::max(1, std::min(64, (std::max(max_send, c->n_ghost) + 1023) / 1024));
			// fused push: every send list one run of consecutive rows; [gap_lo, gap_hi) = the largest run of local indices
			// that nobody needs (interior of a z-slab), so that the pushing kernels test two bounds per access
			d.all_contiguous = 1;
			std::vector<std::pair<int, int>> runs;
			for (const NcclComm::Peer& pe : c->peers)
			{
				if (pe.send_count > 0 && !pe.contiguous) d.all_contiguous = 0;
				if (pe.send_count > 0) runs.emplace_back(pe.send_first, pe.send_first + pe.send_count);
			}
			std::sort(runs.begin(), runs.end());
			int best_lo = 0, best_hi = 0, cur = 0;
			for (const auto& r : runs) { if (r.first - cur > best_hi - best_lo) { best_lo = cur; best_hi = r.first; } cur = std::max(cur, r.second); }
			if (h->n_rows - cur > best_hi - best_lo) { best_lo = cur; best_hi = h->n_rows; }
			d.gap_lo = best_lo; d.gap_hi = best_hi;
			LCG_CUDA_CHECK(cudaDeviceSynchronize());
		}
		// interior / boundary split of the SpMV tiles: boundary tiles (rows that reference ghost columns) go to the end of the
		// tile list, so that k_spmv starts on interior rows while the neighbours' halo entries are still in flight
		h->n_interior = -1;
		if (h->n_tiles > 0 && c->n_ghost > 0)
		{
			int* d_flag = nullptr;
			LCG_CUDA_CHECK(cudaMalloc((void**)&d_flag, sizeof(int) * (size_t)h->n_tiles));
			k_tile_boundary<<<(h->n_tiles * 32 + 255) / 256, 256>>>(h->n_tiles, h->tiles, h->row_ptr, h->col, h->n_rows, d_flag);
			std::vector<int> flag((size_t)h->n_tiles);
			std::vector<int4> tiles((size_t)h->n_tiles), sorted;
			cudaError_t e1 = cudaMemcpy(flag.data(), d_flag, sizeof(int) * flag.size(), cudaMemcpyDeviceToHost);
			cudaError_t e2 = cudaMemcpy(tiles.data(), h->tiles, sizeof(int4) * tiles.size(), cudaMemcpyDeviceToHost);
			cudaFree(d_flag);
			LCG_CUDA_CHECK(e1); LCG_CUDA_CHECK(e2);
			sorted.reserve(tiles.size());
			for (size_t t = 0; t < tiles.size(); t++) if (!flag[t]) sorted.push_back(tiles[t]);
			const int n_int = (int)sorted.size();
			for (size_t t = 0; t < tiles.size(); t++) if (flag[t]) sorted.push_back(tiles[t]);
			LCG_CUDA_CHECK(cudaMemcpy(h->tiles, sorted.data(), sizeof(int4) * sorted.size(), cudaMemcpyHostToDevice));
			h->n_interior = n_int;
		}
		else if (h->n_tiles > 0) h->n_interior = h->n_tiles;
		return 0;
	});
}

int lcgb200_comm_p2p_handle(lcgb200_comm_t comm, void* handle_out, long long* n_ghost_out)
{
	NcclComm* c = reinterpret_cast<NcclComm*>(comm);
	if (!c || !handle_out) return LCGB200_INVALID_POINTER;
	if (!c->window) { set_error_msg("attach a partition (lcgb200_csr_set_partition) before asking for the window handle"); return LCGB200_SIZE_NOT_MATCH; }
	return guarded_comm([&]() {
		cudaIpcMemHandle_t hdl;
		LCG_CUDA_CHECK(cudaIpcGetMemHandle(&hdl, c->window));
		static_assert(sizeof(hdl) == LCGB200_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
		std::memcpy(handle_out, &hdl, sizeof(hdl));
		if (n_ghost_out) *n_ghost_out = c->n_ghost;
		return 0;
	});
}

int lcgb200_comm_p2p_detach(lcgb200_comm_t comm)
{	// back to the NCCL transport (e.g. when some rank could not map its peers' windows)
	NcclComm* c = reinterpret_cast<NcclComm*>(comm);
	if (!c) return LCGB200_INVALID_POINTER;
	c->p2p_ready = false;
	return 0;
}

int lcgb200_comm_p2p_attach(lcgb200_comm_t comm, const void* handles, const long long* n_ghost_of_rank, const long long* remote_off)
{
	NcclComm* c = reinterpret_cast<NcclComm*>(comm);
	if (!c || !handles || !n_ghost_of_rank || !remote_off) return LCGB200_INVALID_POINTER;
	if (!c->window || !c->d_dev) return LCGB200_SIZE_NOT_MATCH;
	return guarded_comm([&]() {
		CommDev& d = c->h_dev;
		for (int r = 0; r < c->size_; r++)
		{
			if (r == c->rank_) { d.win[r] = c->window; continue; }
			cudaIpcMemHandle_t hdl;
			std::memcpy(&hdl, static_cast<const char*>(handles) + (size_t)r * sizeof(hdl), sizeof(hdl));
			void* mapped = nullptr;
			LCG_CUDA_CHECK(cudaIpcOpenMemHandle(&mapped, hdl, cudaIpcMemLazyEnablePeerAccess));
			c->peer_map[r] = mapped;
			d.win[r] = static_cast<CommWindow*>(mapped);
		}
		for (int p = 0; p < d.n_peers; p++)
		{
			d.remote_off[p] = remote_off[p];
			d.remote_ghost[p] = n_ghost_of_rank[d.peer_rank[p]];
		}
		LCG_CUDA_CHECK(cudaMemcpy(c->d_dev, &d, sizeof(d), cudaMemcpyHostToDevice));
		LCG_CUDA_CHECK(cudaDeviceSynchronize());
		c->p2p_ready = true;
		return 0;
	});
}

}  // extern "C"
