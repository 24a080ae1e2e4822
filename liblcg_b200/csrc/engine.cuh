// engine.cuh — host side of a solve: workspace, device state, kernel launch helpers, operator dispatch
// (built-in CSR vs user callback) and the host loop that mirrors the reference's control flow
// (lcg.cpp:206-230: Pfp -> converged? -> max iterations? -> t++ -> work) while keeping the scalars on the device.
#pragma once
#include <functional>
#include <vector>
#include <map>
#include <chrono>
#include <cstdlib>
#include "common.cuh"
#include "csr.cuh"
#include "fused_small.cuh"
#include "ic0.cuh"
#include "exact.cuh"

namespace lcgb200 {

// ---- multi-GPU hooks (comm.cu) ---------------------------------------------------------------------------
struct Comm {
	virtual ~Comm() {}
	virtual int rank() const = 0;
	virtual int size() const = 0;
	// sum-allreduce `count` doubles in place on the device, ordered on `s`
	virtual void allreduce(double* dev, int count, cudaStream_t s) = 0;
	// make the ghost entries of an extended vector (elements [n_local, n_local + n_ghost)) available: NCCL transport
	// fills the vector's ghost tail; the NVLink transport (p2p) pushes my boundary entries into the peers' mailboxes
	virtual void halo(void* x_ext, int elem_bytes, cudaStream_t s, bool p2p, DevState* st) = 0;
	// device state of the NVLink peer-memory transport, or null when only NCCL is available
	virtual CommDev* dev() { return nullptr; }
	// NVLink transport, push half only: my boundary entries of x go into the neighbours' mailboxes (the receive half is in
	// k_spmv's boundary tiles)
	virtual void push(const void* x, int elem_bytes, cudaStream_t s, DevState* st) = 0;
	// every send list is a run of consecutive rows: the kernels that write SpMV inputs can push while they write
	virtual bool fused_push_ok() const { return false; }
	// a cross-GPU wait of an earlier solve timed out: the transport state is out of step with the peers
	virtual bool poisoned() const { return false; }
	virtual bool check_abort() { return false; }
};

// ---- handle behind lcgb200_csr_t ------------------------------------------------------------------------
struct CsrHandle {
	int value_type = 0;            // LCGB200_REAL / LCGB200_COMPLEX
	int n_rows = 0, n_cols = 0, nnz = 0;
	long long n_global = 0;        // rows of the whole (possibly partitioned) system
	unsigned flags = 0;
	int device = 0;
	// device arrays (owned)
	int* row_ptr = nullptr; int* col = nullptr; void* val = nullptr; int4* tiles = nullptr;
	int n_tiles = 0, lpr = 1, chunk = 1;
	int n_interior = -1;           // partitioned row block: tiles [n_interior, n_tiles) reference ghost columns (set by lcgb200_csr_set_partition)
	// transpose (optional)
	int* t_row_ptr = nullptr; int* t_col = nullptr; void* t_val = nullptr; int4* t_tiles = nullptr;
	int t_n_tiles = 0, t_lpr = 1, t_chunk = 1;
	void* diag = nullptr;          // Jacobi diagonal (optional), n_rows values
	// dictionary-compressed copy (LCGB200_CSR_COMPRESS, real): 16-bit codes + dictionaries + its own (larger) tiles
	unsigned short* code = nullptr; double* vdict = nullptr; int* odict = nullptr; int4* dtiles = nullptr;
	int n_dtiles = 0, dchunk = 1, dlpr = 1, n_vdict = 0, n_odict = 0;
	// row-pattern copy: one id per row + the table of distinct rows
	unsigned char* pat = nullptr; unsigned char* pat_thread = nullptr; void* pat_chain = nullptr;
	unsigned char* pat_bitem = nullptr; void* pat_segs = nullptr; PatMarch* pat_march = nullptr;   // plane-marching kernel (optional)
	PatBox* pat_box = nullptr; unsigned char* pat_box_flags = nullptr;                              // box kernel (dense box stencils)
	int n_pat = 0, pat_maxch = 0, pat_stride = 0, pat_nib = 0, pat_items = 0;
	// IC(0) preconditioner (LCGB200_CSR_IC0): L (rows, diagonal last) and U = L^T (rows, diagonal first) with their level
	// orders, plus the vector between the two triangular solves
	IcDev icL, icU; void* ic_tmp = nullptr; bool has_ic0 = false;
	void* user = nullptr;          // instance handed to progress callbacks
	Comm* comm = nullptr;          // set for a row block of a partitioned system
	// partitioned systems: the rows [r0, r1) of A^T live in a second partitioned handle with its own halo plan and windows
	// (complex BiCG's A^H d2, clcg.cpp:188); not owned.  row_offset = global index of the first local row (the shadow
	// residual of complex CGS/BICGSTAB/TFQMR is one rand() sequence over the WHOLE vector, lcg_complex.cpp:118-127)
	const CsrHandle* t_handle = nullptr;
	long long row_offset = 0;
	// cached workspace
	void* ws = nullptr; size_t ws_bytes = 0;
	DevState* d_state = nullptr; DevState* h_state = nullptr; DevState* h_state2 = nullptr; double* d_partials = nullptr;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	// solves on the built-in operator that were handed the legacy default stream run on this stream instead (ordered after
	// the caller's stream by an event; the entry points are host-synchronous): a private stream can be captured into CUDA graphs
	cudaStream_t own_stream = nullptr; cudaEvent_t own_event = nullptr;

	CommDev* p2p_dev() const { return comm ? comm->dev() : nullptr; }
	// NVLink transport with the receive half of the halo exchange inside k_spmv (boundary tiles wait for the neighbours'
	// pushes and read the mailbox in place); the compressed operator copies keep the stand-alone exchange kernel
	bool halo_in_spmv() const { return comm && comm->size() > 1 && comm->dev() && n_interior >= 0 && !code && !pat && !legacy_halo(); }
	static bool legacy_halo() { static const bool v = getenv("LCGB200_HALO_LEGACY") != nullptr; return v; }
	template <class T> CsrDev<T> view() const
	{
		CsrDev<T> v; v.n_rows = n_rows; v.n_cols = n_cols; v.nnz = nnz; v.n_tiles = n_tiles; v.lpr = lpr; v.chunk = chunk;
		v.n_interior = halo_in_spmv() ? n_interior : -1;
		v.comm = halo_in_spmv() ? comm->dev() : nullptr;
		v.row_ptr = row_ptr; v.col = col; v.val = (const T*)val; v.tiles = tiles;
		v.code = code; v.vdict = vdict; v.odict = odict; v.dtiles = dtiles; v.n_dtiles = n_dtiles; v.dchunk = dchunk; v.dlpr = dlpr;
		v.pat = pat; v.pat_thread = pat_thread; v.pat_chain = pat_chain; v.pat_bitem = pat_bitem; v.pat_segs = pat_segs; v.pat_march = pat_march; v.pat_box = pat_box; v.pat_box_flags = pat_box_flags;
		v.n_pat = n_pat; v.pat_maxch = pat_maxch; v.pat_stride = pat_stride; v.pat_nib = pat_nib; v.pat_items = pat_items;
		return v;
	}
	template <class T> CsrDev<T> tview() const
	{
		CsrDev<T> v; v.n_rows = n_cols; v.n_cols = n_rows; v.nnz = nnz; v.n_tiles = t_n_tiles; v.lpr = t_lpr; v.chunk = t_chunk;
		v.row_ptr = t_row_ptr; v.col = t_col; v.val = (const T*)t_val; v.tiles = t_tiles; return v;
	}
};

// user-callback operator: y = op(A) x on raw device pointers (the API layer wraps descriptors around them)
template <class T> using ApplyFn = std::function<void(const T* x, T* y, int op)>;
// progress: (residual, k) -> nonzero to stop
using ProgressFn = std::function<int(double, int)>;

template <class T>
struct Operator {
	const CsrHandle* h = nullptr;   // built-in when non-null
	ApplyFn<T> apply;               // otherwise
	ApplyFn<T> precond;             // user M^-1 (generic path); empty when built-in Jacobi or none
	const T* diag = nullptr;        // built-in Jacobi diagonal
	const CsrHandle* ic0 = nullptr; // built-in IC(0): z = L^-T L^-1 r by two level-ordered triangular solves (ic0.cuh)
	bool host_side = false;         // the callbacks run on the HOST (lcg.h API): check convergence before every call, never run ahead
};

struct Settings {
	long shadow_seed = 0;
	int cres_mode = 0;
	int poll = 4;
	int fused_small = 1;           // cache-resident systems: several whole iterations per cooperative launch (fused_small.cuh)
	int profile = 0;               // 1: bracket every kernel launch with CUDA events (bench roofline pass)
	long long spin_timeout_ms = -1; // cross-GPU waits (NVLink transport); -1 = LCGB200_SPIN_TIMEOUT_MS or 30 s; 0 = wait for ever
	int graphs = -1;               // CUDA graph per batch of iterations; -1 = LCGB200_GRAPHS or automatic
	int pdl = -1;                  // programmatic dependent launch between the kernels of an iteration; -1 = LCGB200_PDL or on
	int fuse_vec2 = -1;            // update + direction in one cooperative kernel; -1 = LCGB200_FUSE_VEC2 or automatic (by size)
	int l2_persist = -1;           // persisting-L2 window over the work vectors; -1 = LCGB200_L2_PERSIST or automatic
	int reference_order = -1;      // reference-order arithmetic (exact.cuh): bit-identical to the reference's CPU build; -1 = LCGB200_REFERENCE_ORDER or off
};
bool reference_order();
long long spin_timeout_ms();
// cooperative launch that also carries the programmatic-stream-serialisation attribute when PDL is active for this solve
// (the kernel starts with pdl_enter()); falls back to the plain cooperative launch if the driver refuses the combination
cudaError_t launch_coop_pdl(const void* kern, int grid, int block, void** args, cudaStream_t s);
bool fuse_vec2(size_t n_local);
int coop_grid_full(const void* kernel, int block);   // all co-resident blocks of a kernel on the current device (engine.cu)
Settings& settings();

class Engine {
public:
	cudaStream_t stream = nullptr;
	DevState* d_st = nullptr;
	DevState* h_st = nullptr;     // pinned
	DevState* h_st2 = nullptr;    // pinned (second poll slot)
	double* d_partials = nullptr;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	Comm* comm = nullptr;
	bool own_state = false;
	int launches = 0, spmv_launches = 0;
	// profile mode: event pairs around launches, class 0 = SpMV (+fused dots), 1 = fused vector kernels
	struct Timed { cudaEvent_t a, b; int cls; };
	std::vector<Timed> timed; size_t timed_used = 0;
	bool profiling = false;
	cudaEvent_t prof_begin(int cls);
	void prof_end(cudaEvent_t e) { if (e) cudaEventRecord(e, stream); }
	void prof_collect(double* ms, int* count);   // sums per class (2 entries each); call after a stream sync
	ProgressFn pf;                 // empty = no progress callback
	bool sync_each = false;        // one host round trip per loop head even without a progress callback (user operator callbacks:
	                               // the reference never calls Afp / Mfp after convergence, so the host must not run ahead of the test)
	bool capturable = true;        // the iteration is a fixed list of kernel launches (no host-visible sync point inside): a batch may be a CUDA graph
	int seen_checks = 0;
	int final_ret = RC_UNKNOWN;
	// workspace arena
	char* ws = nullptr; size_t ws_cap = 0, ws_off = 0; bool own_ws = false;
	CsrHandle* cache = nullptr;    // handle whose cached buffers we borrow

	Engine(cudaStream_t s, CsrHandle* h);
	~Engine();

	void reserve(size_t bytes);
	size_t ws_need = 0;            // bytes of the arena this solve uses
	size_t l2_from = 0;            // arena offset where the iteration's vectors start (copies of B / the box come first)
	size_t l2_unit = 0;            // bytes of one work vector in the arena
	// L2 residency of the work vectors (mid-size systems whose vectors fit the 126 MB L2 while the matrix does not): the
	// arena becomes a persisting access-policy window of the solve's stream for the duration of the solve
	bool l2_window_set = false;
	size_t l2_prev_limit = 0;      // the device's persisting carve-out before this solve changed it
	void l2_window_begin();
	void l2_window_end();
	std::vector<void*> allocs;     // every vector handed out, in order (lcg()/lcgs() copy their work vectors back to the caller)
	template <class T> T* alloc(size_t count)
	{
		size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
		if (ws_off + bytes > ws_cap) { set_error_msg("workspace overflow"); throw CudaFailure(); }
		T* p = reinterpret_cast<T*>(ws + ws_off); ws_off += bytes; allocs.push_back(p); return p;
	}

	void start(const DevState& init);   // upload the initial state, record the start event
	void read_state();                  // D2H of the state block + stream sync
	double device_ms();                 // start..now on the stream (syncs)

	bool multi() const { return comm != nullptr && comm->size() > 1; }
	bool p2p() const { return multi() && cache && cache->p2p_dev() != nullptr; }

	template <class Op, class V = Variant> void vec(const Op& op, size_t n)
	{
		if constexpr (V::exact)
		{	// reference-order build: terms in parallel, serial totals + scalar epilogue in a second launch
			launch_exact_vec(op, n, d_st, d_terms, terms_stride, stream);
			launches += Op::NRED > 0 ? 2 : 1;
		}
		else
		{
			cudaEvent_t pe = profiling ? prof_begin(1) : nullptr;
			launch_k(k_vec<Op, false, double>, vec_grid(n, Op::W), kThreads, 0, stream, op, n, d_st, d_partials, (CommDev*)nullptr, (const double*)nullptr);
			prof_end(pe);
			launches++;
			if (Op::NRED > 0 && multi()) finish_multi(op, Op::NRED);
		}
	}

	// Same, for a kernel that WRITES `out`, the input of the next SpMV.  On the NVLink transport (row partition whose send
	// lists are runs of consecutive rows) the kernel pushes the entries the neighbours need into their mailboxes while it
	// writes them, and the SpMV that follows needs no exchange launch: its interior tiles start at once, its boundary tiles
	// wait for the neighbours' flags (halo overlapped with the interior rows).
	const void* pushed_vec = nullptr;   // the vector whose halo the last pushing kernel has already sent
	bool halo_in_spmv() const { return multi() && cache && cache->halo_in_spmv(); }
	template <class T, class Op, class V = Variant> void vec_push(const Op& op, size_t n, const T* out)
	{
		if constexpr (V::exact) vec(op, n);
		else
		{
			if (!(halo_in_spmv() && comm->fused_push_ok())) { vec(op, n); return; }
			cudaEvent_t pe = profiling ? prof_begin(1) : nullptr;
			launch_k(k_vec<Op, true, T>, vec_grid(n, Op::W), kThreads, 0, stream, op, n, d_st, d_partials, cache->p2p_dev(), out);
			prof_end(pe);
			launches++;
			pushed_vec = out;
			if (Op::NRED > 0 && multi()) finish_multi(op, Op::NRED);
		}
	}

	// update + direction in one cooperative launch (k_vec2) where that is possible: our own reductions (single GPU or the
	// NVLink transport: the NCCL transport needs a collective between the two steps), a grid that is wholly co-resident, and
	// systems small enough for a kernel boundary to matter (the regime of the CUDA graphs); otherwise two launches.
	template <class T, class Op1, class Op2, class V = Variant> void vec2_push(const Op1& a, const Op2& b, size_t n, const T* out)
	{
		if constexpr (V::exact) { vec(a, n); vec(b, n); }
		else
		{
			if (!fuse_vec2(n) || (multi() && !p2p())) { vec(a, n); vec_push(b, n, out); return; }
			const bool push = halo_in_spmv() && comm->fused_push_ok();
			CommDev* cd = push ? cache->p2p_dev() : nullptr;
			DevState* st = d_st; double* parts = d_partials; size_t nn = n;
			Op1 a1 = a; Op2 b1 = b;
			void* args[] = {&a1, &b1, &nn, &st, &parts, &cd, (void*)&out};
			const void* kern = push ? (const void*)k_vec2<Op1, Op2, true, T> : (const void*)k_vec2<Op1, Op2, false, T>;
			int grid = vec_grid(n, Op1::W);
			const int limit = coop_grid_full(kern, kThreads);
			if (grid > limit) grid = limit;
			cudaEvent_t pe = profiling ? prof_begin(1) : nullptr;
			LCG_CUDA_CHECK(launch_coop_pdl(kern, grid, kThreads, args, stream));
			prof_end(pe);
			launches++;
			if (push) pushed_vec = out;
		}
	}

	template <class Op> void finish_multi(const Op& op, int nred)
	{
		if (p2p()) return;   // NVLink transport: the producing kernel's last block already summed across the ranks and ran the epilogue
		comm->allreduce(reinterpret_cast<double*>(reinterpret_cast<char*>(d_st) + offsetof(DevState, red)), nred, stream);
		launch_k(k_finish<Op>, 1, 1, 0, stream, op, d_st);
		launches++;
	}

	// y = op(A) x with a fused per-row epilogue.  op: 0 = N, 1 = T, 2 = H.
	template <class T, class Epi, class V = Variant> void spmv(const Operator<T>& A, T* x, T* y, const Epi& epi, int op = 0)
	{
		if constexpr (V::exact)
		{
			if (A.h)
			{	// built-in operator, reference order: one thread per row over the plain CSR arrays (A^T / A^H: the stored transpose)
				const CsrDev<T> Vw = op == 0 ? A.h->template view<T>() : A.h->template tview<T>();
				if (op == 2) launch_exact_spmv<T, true, Epi>(Vw, x, y, epi, d_st, d_terms, terms_stride, stream);
				else launch_exact_spmv<T, false, Epi>(Vw, x, y, epi, d_st, d_terms, terms_stride, stream);
				launches += Epi::NRED > 0 ? 2 : 1; spmv_launches++;
			}
			else
			{
				A.apply(x, y, op);
				spmv_launches++;
				if (Epi::ACTIVE) vec(RowEpilogueOp<T, Epi>{epi, x, y}, (size_t)n_local);
			}
		}
		else if (A.h)
		{
			// op(A) of a partitioned system: A^T / A^H come from the second partitioned handle (its own plan, windows, sequence)
			const CsrHandle* hh = (op != 0 && A.h->t_handle) ? A.h->t_handle : A.h;
			if (hh->comm && hh->comm->size() > 1)
			{
				if (hh->halo_in_spmv())
				{	// receive half inside k_spmv; push half only if the producing kernel has not done it already
					if (hh != A.h || pushed_vec != (const void*)x) { hh->comm->push(x, (int)sizeof(T), stream, d_st); launches++; }
					if (hh == A.h) pushed_vec = nullptr;
				}
				else { hh->comm->halo(x, (int)sizeof(T), stream, hh->p2p_dev() != nullptr, d_st); if (hh->p2p_dev()) launches++; }
			}
			const CsrDev<T> Vw = (hh != A.h || op == 0) ? hh->template view<T>() : hh->template tview<T>();
			cudaEvent_t pe = profiling ? prof_begin(0) : nullptr;
			if (op == 2) launch_spmv<T, true, Epi>(Vw, x, y, epi, d_st, d_partials, stream);
			else launch_spmv<T, false, Epi>(Vw, x, y, epi, d_st, d_partials, stream);
			prof_end(pe);
			launches++; spmv_launches++;
			if (Epi::NRED > 0 && multi()) finish_multi(RowEpilogueOp<T, Epi>{epi, x, y}, Epi::NRED);
		}
		else
		{
			A.apply(x, y, op);
			spmv_launches++;
			if (Epi::ACTIVE) vec(RowEpilogueOp<T, Epi>{epi, x, y}, (size_t)n_local);
		}
	}

	size_t n_local = 0;
	// reference-order build: one term per element and reduction slot (kMaxRed x terms_stride doubles, from the workspace arena)
	double* d_terms = nullptr; size_t terms_stride = 0;

	// z = (L L^T)^-1 r with the handle's IC(0) factor: forward solve into the handle's scratch vector, backward solve into z
	template <class T, class V = Variant> void ic0_solve(const CsrHandle* h, const T* r, T* z)
	{
		if constexpr (V::exact) { set_error_msg("the IC(0) preconditioner is not available in reference-order mode"); throw CudaFailure(); }
		else
		{
			cudaEvent_t pe = profiling ? prof_begin(1) : nullptr;
			launch_sptrsv<T, false>(h->icL, r, static_cast<T*>(h->ic_tmp), d_st, stream);
			launch_sptrsv<T, true>(h->icU, static_cast<const T*>(h->ic_tmp), z, d_st, stream);
			prof_end(pe);
			launches += 2;
		}
	}
	// M^-1 of the preconditioned solvers' generic path: the built-in IC(0) or the user's callback
	template <class T, class V = Variant> void precondition(const Operator<T>& A, const T* r, T* z)
	{
		if (A.ic0) ic0_solve<T, V>(A.ic0, r, z);
		else A.precond(r, z, 0);
	}

	// cache-resident single-GPU systems on the built-in operator take the fused cooperative kernel (not while profiling:
	// the per-kernel attribution of bench.py's roofline pass needs the separate launches)
	template <class T, class V = Variant> bool small_system(const Operator<T>& A) const
	{
		if (V::exact) return false;
		return A.h && !multi() && !profiling && settings().fused_small && A.h->n_rows <= kSmallRows && A.h->nnz <= kSmallNnz;
	}
	// phase builders for the fused kernel (what E.spmv / E.vec would have launched)
	template <class T, class Epi> SpmvPhase<T, false, Epi> ph_spmv(const Operator<T>& A, const T* x, T* y, const Epi& epi) const
	{
		return SpmvPhase<T, false, Epi>{A.h->template view<T>(), x, y, epi};
	}
	template <class T, class Epi> SpmvPhase<T, true, Epi> ph_spmv_h(const Operator<T>& A, const T* x, T* y, const Epi& epi) const   // A^H x
	{
		return SpmvPhase<T, true, Epi>{A.h->template tview<T>(), x, y, epi};
	}
	template <class Op> VecPhase<Op> ph_vec(const Op& op, size_t n) const { return VecPhase<Op>{op, n}; }
	// `iters` iterations of the phase list in one cooperative launch; n_spmv = SpMV phases per iteration
	template <class... Ph> void fused(int iters, int n_spmv, Ph... ph)
	{
		if constexpr (Variant::exact) { set_error_msg("no fused path in the reference-order build"); throw CudaFailure(); }
		else
		{
			LCG_CUDA_CHECK((launch_fused<Ph...>(d_st, d_partials, iters, stream, ph...)));
			launches++; spmv_launches += iters * n_spmv;
		}
	}

	// Pfp mode: read the state, deliver a new loop head to the callback.  Returns true when the solve is over
	// (final_ret set).  Without a callback this is a no-op returning false (the device flags do the work).
	bool sync_point();
	// unconditional state read (SPG line search); returns true when the solve is over
	bool sync_always();

	// Host loop.  `iterate` enqueues one iteration and returns true if a sync point inside it ended the solve.
	// `batch`, when given, enqueues k iterations in one go (fused cooperative kernel) and replaces `iterate`.
	int run(const std::function<bool()>& iterate, const std::function<void(int)>& batch = nullptr);
};

}  // namespace lcgb200
