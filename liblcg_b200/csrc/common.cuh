// common.cuh — device state block, deterministic single-pass grid reduction, vector access helpers and the
// generic fused "vector update + reductions + scalar epilogue" kernel every solver step is built from.
//
// Design: all iteration scalars (alpha, beta, omega, the squared norms, the iteration counter t,
// the return code) live in ONE device-resident struct.  The last block of each kernel finishes the reduction in
// a fixed order (run-to-run deterministic) and runs the step's scalar epilogue — including the reference's
// loop-head control (lcg.cpp:206-230): residual, convergence test, max-iteration test, t++.  Once `done` is set
// every later kernel of the solve returns immediately, so the host may enqueue iterations ahead of the
// convergence test without ever applying an extra update to the solution.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>

namespace lcgb200 {

// ---- arithmetic variants ------------------------------------------------------------------------------------
// The solver translation units are compiled twice.  The default build (VariantStd) is the fast path: fused
// multiply-adds and tree-ordered reductions.  The second build (-DLCG_REFORDER -fmad=false, VariantExact) reproduces the
// arithmetic of the reference's x86-64 CPU build operation for operation — every a*b+c rounds twice (gcc emits no FMA
// for the reference's flags, src/CMakeLists.txt:39-40), every row sum and every dot product adds its terms left to right
// in index order (algebra.cpp:154-163, lcg_complex.cpp:143-167) — so that iterates, residual histories and iteration
// counts are bit-identical to the reference's (lcgb200_set_reference_order; exact.cuh).  The variant is a (defaulted)
// template argument of every Engine launch helper: the two builds instantiate differently named functions and kernels.
struct VariantStd { static constexpr bool exact = false; };
struct VariantExact { static constexpr bool exact = true; };
#ifdef LCG_REFORDER
typedef VariantExact Variant;
__device__ __forceinline__ double lcg_mul_add(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }
__device__ __forceinline__ float lcg_mul_addf(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }
#define fma(a, b, c) lcg_mul_add((a), (b), (c))
#define fmaf(a, b, c) lcg_mul_addf((a), (b), (c))
#else
typedef VariantStd Variant;
#endif

constexpr int kThreads = 256;        // threads per block, all kernels
constexpr int kMaxBlocks = 148 * 8;  // persistent grids: multiples of the 148 SMs of a B200
constexpr int kMaxRed = 8;           // max reduction slots per kernel
constexpr int kMaxWarps = 16;        // max warps per block of any kernel that uses grid_reduce
constexpr int kNumSc = 40;

// return codes (reference util.h:69-90)
enum : int {
	RC_CONVERGENCE = 0, RC_STOP = 1, RC_ALREADY = 2, RC_UNKNOWN = -1024, RC_BAD_SIZE = -1023, RC_BAD_MAXIT = -1022,
	RC_BAD_EPS = -1021, RC_BAD_RESTART = -1020, RC_MAXIT = -1019, RC_NULL_PRECOND = -1018, RC_NAN = -1017,
	RC_BAD_PTR = -1016, RC_BAD_LAMBDA = -1015, RC_BAD_SIGMA = -1014, RC_BAD_BETA = -1013, RC_BAD_MAXIM = -1012,
	RC_C_NAN = -1019, RC_C_BAD_PTR = -1018, RC_C_UNKNOWN_SOLVER = -1016
};

// scalar register file of a solve (indices into DevState::sc); complex values take two consecutive slots
enum : int {
	SC_MMOD = 0,     // max(m.m, 1)            (complex: max(|m|^4, 1))
	SC_RMOD,         // squared residual norm  (complex: |r|^4)
	SC_RHO,          // CG: g.g | PCG: z.r | CGS/BICGSTAB: r.r0~   (+1 imag)
	SC_RHO_I,
	SC_ALPHA,        // (+1 imag)
	SC_ALPHA_I,
	SC_BETA,
	SC_BETA_I,
	SC_OMEGA,
	SC_OMEGA_I,
	SC_TMP0, SC_TMP1, SC_TMP2, SC_TMP3, SC_TMP4, SC_TMP5,
	SC_STEP,         // PG alpha_k / SPG lambda_k
	SC_LS_ALPHA,     // SPG line-search alpha
	SC_QK,           // SPG objective of the trial point
	SC_GD,           // SPG g.d
	SC_THETA, SC_TAO, SC_ETA, SC_ETA_I, SC_RKM, SC_RKM2,  // TFQMR
	SC_PART0,        // partial sums carried between the halves of a split (callback-preconditioned) update
	SC_PART1, SC_PART2, SC_PART3
};

// ---- multi-GPU over NVLink peer memory (comm.cu) ------------------------------------------------------------
// Every rank owns one CommWindow in device memory, mapped into all peers with CUDA IPC.  Peers PUSH into it:
//   * the per-rank totals of a fused reduction (ar_val / ar_flag): the last block of the producing kernel stores its
//     totals into every rank's window, so the "allreduce" is 8 remote stores in the kernel's tail plus a local sum;
//   * the halo entries of the next SpMV input (mailbox behind the header, double-buffered) and a sequence flag.
// Flags carry monotonically increasing sequence numbers (never reset), written with st.release.sys after a
// __threadfence_system() and polled with ld.acquire.sys.
constexpr int kMaxRanks = 8;
constexpr int kArSlots = 4;
struct CommWindow {
	// reduction totals, one 64-bit word per HALF double: low 32 bits = data, high 32 bits = sequence tag.  An aligned 8-byte
	// store lands atomically, so a word whose tag matches carries valid data: no fence and no separate flag between the
	// payload and its arrival signal (the "LL" idea of NCCL's low-latency protocol) — one NVLink one-way trip per reduction.
	unsigned long long ar_ll[kArSlots][kMaxRanks][2 * kMaxRed];
	unsigned long long halo_flag[kMaxRanks];     // [source rank]: sequence number of the last halo it pushed to me
	unsigned long long halo_ack[kMaxRanks];      // [receiving rank]: sequence number of the last halo of MINE it has consumed
	// mailbox: 2 buffers x n_ghost x 16 bytes follow (offset kMailboxOffset)
};
constexpr size_t kMailboxOffset = (sizeof(CommWindow) + 255) & ~size_t(255);

struct CommDev {	// device-resident, private to the rank
	unsigned long long ar_seq, halo_seq;
	int rank, size, n_local, n_ghost, n_peers, n_packed;
	unsigned int ticket, ticket2, abort_flag, pad0;
	CommWindow* win[kMaxRanks];                  // win[rank] = own window, others = IPC mappings of the peers' windows
	int peer_rank[kMaxRanks], send_first[kMaxRanks], send_count[kMaxRanks], send_off[kMaxRanks], contiguous[kMaxRanks], recv_count[kMaxRanks];
	long long remote_off[kMaxRanks];             // where my entries start inside the peer's ghost region
	long long remote_ghost[kMaxRanks];           // the peer's n_ghost (size of one of its mailbox buffers)
	const int* send_idx;                         // packed (non-contiguous) send indices
	// fused push (the kernel that WRITES an SpMV input stores the entries its neighbours need straight into their
	// mailboxes): possible when every send list is one run of consecutive rows (z-slabs of a stencil, banded matrices);
	// local indices in [gap_lo, gap_hi) are sent to nobody
	int all_contiguous, gap_lo, gap_hi, pad1;
};

// mailbox buffer `buf` of window w, as elements of type T
template <class T> __device__ __forceinline__ T* mailbox_of(CommWindow* w, unsigned long long buf, long long n_ghost)
{
	return reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(w) + kMailboxOffset) + (size_t)(buf & 1ull) * (size_t)n_ghost;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// flag store AFTER an explicit __threadfence_system(): no second fence (a st.release.sys per peer would pay one NVLink
// round trip each)
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
// spin until *flag >= want; false after timeout_ns (0 = wait for ever).  The timeout is a per-solve setting carried in
// DevState::spin_timeout_ns (lcgb200_set_spin_timeout_ms / LCGB200_SPIN_TIMEOUT_MS; longer when the host is in the loop):
// a peer that never shows up ends the solve with an error instead of a hang, and the communicator is then poisoned.
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long want, unsigned long long timeout_ns)
{
	if (ld_acquire_sys(flag) >= want) return true;
	const unsigned long long t0 = global_ns();
	while (ld_acquire_sys(flag) < want)
	{
		__nanosleep(64);
		if (timeout_ns && global_ns() - t0 > timeout_ns) return false;
	}
	return true;
}

struct DevState {
	double sc[kNumSc];
	double red[kMaxRed];      // totals of the last reduction (multi-GPU: all-reduced in place before finish)
	double residual;          // value handed to the progress callback
	double eps, restart_eps, sigma, ls_beta;
	long long n_global;       // n of the whole system (the abs_diff test divides by it)
	int abs_diff, max_it, cres_mode, multi;
	int t;                    // the reference's iteration counter
	int k_report;             // t at the last loop head (what Pfp receives)
	int done, ret;
	int checks;               // number of loop heads executed (host detects new heads by watching it)
	int flag;                 // solver-specific (BICGSTAB2 half-step convergence, SPG line-search accept)
	int half;                 // TFQMR half-step index / BICGSTAB2 restart marker
	unsigned int ticket;
	unsigned int pad;
	CommDev* comm;            // multi == 2: NVLink peer-memory transport
	unsigned long long spin_timeout_ns;   // cross-GPU waits give up after this long (0 = never)
	int gate;                 // k_vec2: bumped by the block that finished the first step's reduction, releases the second step
	int pad2;
};

// multi-GPU: the cross-rank sum of a fused reduction, called by the whole first warp of the last block (grid_reduce).
//   multi == 1 (NCCL transport): leave the local totals in st->red for ncclAllReduce + k_finish; returns false.
//   multi == 2 (NVLink peer memory): lane p pushes the local totals into rank p's window — all peers in parallel, one
//     system-scope fence, one flag store each, i.e. a single NVLink round trip — then the same warp waits until every
//     rank's totals have landed in MY window and sums them in rank order (bitwise identical on all ranks).  Returns true
//     with tot[] = global totals in lane 0: the producing kernel's tail IS the allreduce, the caller runs the scalar
//     epilogue right there and no separate launch exists.  A peer that never arrives ends the solve with an error.
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ bool reduce_across_ranks(DevState* st, double* tot, int nred)
{
	const int lane = threadIdx.x & 31;
	if (st->multi != 2)
	{
		if (lane == 0) for (int r = 0; r < nred; r++) st->red[r] = tot[r];
		return false;
	}
	CommDev* c = st->comm;
	const unsigned long long seq = c->ar_seq + 1;
	const unsigned long long tag = (seq & 0xffffffffull) << 32;
	const int slot = (int)(seq % kArSlots);
	if (lane < c->size)
	{	// lane p -> rank p's window, all peers in parallel; tagged words need no fence
		unsigned long long* dst = c->win[lane]->ar_ll[slot][c->rank];
		for (int r = 0; r < nred; r++)
		{
			const unsigned long long bits = (unsigned long long)__double_as_longlong(tot[r]);
			st_relaxed_sys(dst + 2 * r, (bits & 0xffffffffull) | tag);
			st_relaxed_sys(dst + 2 * r + 1, (bits >> 32) | tag);
		}
	}
	// lane s collects rank s's totals out of MY window
	double mine[kMaxRed];
	bool ok = true;
	if (lane < c->size)
	{
		const unsigned long long* src = c->win[c->rank]->ar_ll[slot][lane];
		const unsigned long long t0 = global_ns();
		for (int r = 0; r < nred && ok; r++)
		{
			unsigned long long lo, hi;
			while (true)
			{
				lo = ld_relaxed_sys(src + 2 * r); hi = ld_relaxed_sys(src + 2 * r + 1);
				if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
				if (st->spin_timeout_ns && global_ns() - t0 > st->spin_timeout_ns) { ok = false; break; }
			}
			mine[r] = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
		}
	}
	ok = __all_sync(0xffffffffu, ok);
	// every lane sums the ranks' values in rank order: bitwise identical totals on all ranks
	for (int r = 0; r < nred; r++)
	{
		double v = 0.0;
		for (int srk = 0; srk < c->size; srk++) v += __shfl_sync(0xffffffffu, (lane < c->size && ok) ? mine[r] : 0.0, srk);
		tot[r] = v;
	}
	if (lane != 0) return false;
	c->ar_seq = seq;
	if (!ok) { st->ret = RC_UNKNOWN; st->done = 1; c->abort_flag = 1; return false; }
	return true;
}

__device__ __forceinline__ int st_done(const DevState* st) { return *((volatile const int*)&st->done); }

// ---- reference loop head, real solvers (lcg.cpp:206-230) -------------------------------------------------
__device__ __forceinline__ void loop_head_real(DevState* st, double sq_res, double m_mod)
{
	const double residual = st->abs_diff ? sqrt(sq_res) / (double)st->n_global : sq_res / m_mod;
	st->residual = residual;
	st->k_report = st->t;
	st->checks++;
	if (residual <= st->eps) { st->ret = RC_CONVERGENCE; st->done = 1; return; }
	if (st->max_it > 0 && st->t + 1 > st->max_it) { st->ret = RC_MAXIT; st->done = 1; return; }
	st->t++;
}

// "already optimised" test before the loop (lcg.cpp:185-203); falls through into the first loop head
__device__ __forceinline__ void first_head_real(DevState* st, double sq_res, double m_mod)
{
	double r;
	bool hit = false;
	if (st->abs_diff && (r = sqrt(sq_res) / (double)st->n_global) <= st->eps) hit = true;
	else if ((r = sq_res / m_mod) <= st->eps) hit = true;
	if (hit) { st->residual = r; st->k_report = 0; st->checks++; st->ret = RC_ALREADY; st->done = 1; return; }
	loop_head_real(st, sq_res, m_mod);
}

// ---- complex residual (clcg.cpp:112-147 CPU definition, or clcg_cuda.cu:145-176 when cres_mode = 1) -----
// rr = ||r||^2, mm = ||m||^2 (plain real sums).  CPU: rk_square = |<r,r>|^2 = rr^2, m_square = max(mm^2, 1).
__device__ __forceinline__ double cplx_res_abs(const DevState* st, double rr)
{
	if (st->cres_mode == 0) return sqrt(rr * rr) / (double)st->n_global;   // sqrt(rk_square)/n
	return sqrt(rr) / (double)st->n_global;                                // rk_mod/n (nrm2)
}
__device__ __forceinline__ double cplx_res_rel(const DevState* st, double rr, double mm)
{
	if (st->cres_mode == 0) { double msq = mm * mm; if (msq < 1.0) msq = 1.0; return rr * rr / msq; }
	double mn = sqrt(mm); if (mn < 1.0) mn = 1.0;
	double rn = sqrt(rr);
	return rn * rn / (mn * mn);
}

__device__ __forceinline__ void loop_head_cplx(DevState* st, double rr, double mm)
{
	const double residual = st->abs_diff ? cplx_res_abs(st, rr) : cplx_res_rel(st, rr, mm);
	st->residual = residual;
	st->k_report = st->t;
	st->checks++;
	if (residual <= st->eps) { st->ret = RC_CONVERGENCE; st->done = 1; return; }
	if (st->max_it > 0 && st->t + 1 > st->max_it) { st->ret = RC_MAXIT; st->done = 1; return; }
	st->t++;
}

// clcg.cpp:123-141: the relative test is also tried when abs_diff is set and the absolute one fails
__device__ __forceinline__ void first_head_cplx(DevState* st, double rr, double mm)
{
	double r;
	bool hit = false;
	if (st->abs_diff && (r = cplx_res_abs(st, rr)) <= st->eps) hit = true;
	else if ((r = cplx_res_rel(st, rr, mm)) <= st->eps) hit = true;
	if (hit) { st->residual = r; st->k_report = 0; st->checks++; st->ret = RC_ALREADY; st->done = 1; return; }
	loop_head_cplx(st, rr, mm);
}

// ---- complex arithmetic on double2 (x = re, y = im) ------------------------------------------------------
typedef double2 zc;
__device__ __forceinline__ zc zmk(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ zc zadd(zc a, zc b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ zc zsub(zc a, zc b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ zc zmul(zc a, zc b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ zc zconj(zc a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ zc zscale(double s, zc a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ double znorm2(zc a) { return a.x * a.x + a.y * a.y; }
// Smith's algorithm, the same scaling strategy libgcc's __divdc3 uses for finite operands
__device__ __forceinline__ zc zdiv(zc a, zc b)
{
	double ratio, denom, re, im;
	if (fabs(b.x) < fabs(b.y)) { ratio = b.x / b.y; denom = b.x * ratio + b.y; re = (a.x * ratio + a.y) / denom; im = (a.y * ratio - a.x) / denom; }
	else { ratio = b.y / b.x; denom = b.y * ratio + b.x; re = (a.y * ratio + a.x) / denom; im = (a.y - a.x * ratio) / denom; }
	return make_double2(re, im);
}
// Single-precision complex STORAGE (cuComplex vectors and matrix values of clcg_cudaf.h:81-105): 8 bytes in memory, converts
// to and from the double2 the solver arithmetic uses, so the same step functors serve both precisions — they read with
// `Z v = x[i]` and write with `x[i] = v`.  Scalars, dot products and norms stay in double.
struct __align__(8) ZF {
	float x, y;
	ZF() = default;
	__host__ __device__ ZF(float re, float im) : x(re), y(im) {}
	__host__ __device__ ZF(const double2& v) : x((float)v.x), y((float)v.y) {}
	__host__ __device__ operator double2() const { return make_double2((double)x, (double)y); }
};

__device__ __forceinline__ zc sc_ldz(const DevState* st, int i) { return make_double2(st->sc[i], st->sc[i + 1]); }
__device__ __forceinline__ void sc_stz(DevState* st, int i, zc v) { st->sc[i] = v.x; st->sc[i + 1] = v.y; }

// ---- streaming global access ------------------------------------------------------------------------------
// matrix streams (read once): bypass L1 allocation so L1 stays available for the gathered x
__device__ __forceinline__ int4 ldg_stream_i4(const int4* p)
{
	int4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ double2 ldg_stream_d2(const double2* p)
{
	double2 r;
	asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
	return r;
}
__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }

// W consecutive doubles as one access (W = 2 -> one 128-bit transaction)
template <int W> struct DV { double v[W]; };
template <int W> __device__ __forceinline__ DV<W> dv_load(const double* p);
template <> __device__ __forceinline__ DV<1> dv_load<1>(const double* p) { DV<1> r; r.v[0] = *p; return r; }
template <> __device__ __forceinline__ DV<2> dv_load<2>(const double* p)
{
	double2 t = *reinterpret_cast<const double2*>(p); DV<2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <int W> __device__ __forceinline__ void dv_store(double* p, const DV<W>& a);
template <> __device__ __forceinline__ void dv_store<1>(double* p, const DV<1>& a) { *p = a.v[0]; }
template <> __device__ __forceinline__ void dv_store<2>(double* p, const DV<2>& a)
{
	*reinterpret_cast<double2*>(p) = make_double2(a.v[0], a.v[1]);
}

// ---- deterministic grid reduction -------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// Block-reduce acc[0..NRED), publish per-block partials, elect the last block, and let it total the partials
// in a fixed order (any block size that is a multiple of 32, up to kMaxWarps warps).  Returns true in the 32 threads of
// the first warp of the last block (and nowhere else), each with tot[] holding the grid totals: lane 0 runs the scalar
// epilogue, the whole warp takes part in the cross-GPU publish.  partials must hold gridDim.x * NRED doubles.
template <int NRED>
__device__ __forceinline__ bool grid_reduce(const double* acc, double* partials, unsigned int* ticket, double* tot)
{
	__shared__ double s_red[kMaxRed][kMaxWarps];
	__shared__ int s_last;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (int)(blockDim.x >> 5);
#pragma unroll
	for (int r = 0; r < NRED; r++)
	{
		double v = warp_sum(acc[r]);
		if (lane == 0) s_red[r][warp] = v;
	}
	__syncthreads();
	if (warp == 0)
	{
#pragma unroll
		for (int r = 0; r < NRED; r++)
		{
			double v = (lane < nwarps) ? s_red[r][lane] : 0.0;
			v = warp_sum(v);
			if (lane == 0) partials[(size_t)blockIdx.x * NRED + r] = v;
		}
	}
	if (threadIdx.x == 0)
	{
		__threadfence();
		unsigned int prev = atomicAdd(ticket, 1u);
		s_last = (prev == gridDim.x - 1) ? 1 : 0;
	}
	__syncthreads();
	if (!s_last) return false;
	__threadfence();
#pragma unroll
	for (int r = 0; r < NRED; r++)
	{
		double v = 0.0;
		for (int b = threadIdx.x; b < (int)gridDim.x; b += (int)blockDim.x) v += ld_cg(partials + (size_t)b * NRED + r);
		v = warp_sum(v);
		if (lane == 0) s_red[r][warp] = v;
	}
	__syncthreads();
	if (warp != 0) return false;
#pragma unroll
	for (int r = 0; r < NRED; r++)
	{	// every lane sums the same values in the same order: identical totals without a broadcast
		double v = 0.0;
		for (int w = 0; w < nwarps; w++) v += s_red[r][w];
		tot[r] = v;
	}
	if (lane == 0) *ticket = 0u;
	return true;
}

// ---- fused halo push (multi-GPU, NVLink transport) -----------------------------------------------------------
// Called by the thread that has just written src[i .. i+V): the entries that fall into a neighbour's send range go
// straight into that neighbour's mailbox (remote stores over NVLink).  Returns true if anything was pushed.
template <class T, int V>
__device__ __forceinline__ bool push_elems(const CommDev* c, const T* src, size_t i, unsigned long long seq)
{
	if ((long long)i >= c->gap_lo && (long long)(i + V) <= c->gap_hi) return false;
	bool pushed = false;
	for (int p = 0; p < c->n_peers; p++)
	{
		const long long first = c->send_first[p], cnt = c->send_count[p];
#pragma unroll
		for (int k = 0; k < V; k++)
		{
			const long long e = (long long)i + k - first;
			if (e >= 0 && e < cnt)
			{
				T* dst = mailbox_of<T>(c->win[c->peer_rank[p]], seq, c->remote_ghost[p]) + c->remote_off[p];
				dst[e] = src[i + k];
				pushed = true;
			}
		}
	}
	return pushed;
}

// End of a pushing kernel: every block fences and takes a ticket; the last one bumps my sequence flag in the
// neighbours' windows and publishes the new halo_seq for the consumer (the next SpMV on this stream).
__device__ __forceinline__ void push_signal(CommDev* c, unsigned long long seq, bool pushed)
{
	if (pushed) __threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0)
	{
		__threadfence();
		if (atomicAdd(&c->ticket, 1u) == gridDim.x - 1)
		{
			__threadfence_system();
			for (int p = 0; p < c->n_peers; p++)
				if (c->send_count[p] > 0) st_relaxed_sys(&c->win[c->peer_rank[p]]->halo_flag[c->rank], seq);
			c->halo_seq = seq;
			c->ticket = 0u;
		}
	}
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// The kernels of an iteration form a chain in which each one needs everything its predecessor wrote.  Launched with the
// programmatic-stream-serialization attribute, a kernel may be SCHEDULED while its predecessor is still running: its blocks
// become resident as the predecessor's blocks retire and park in griddepcontrol.wait until the predecessor has completed
// and its writes are visible — the launch latency and block scheduling of kernel k+1 overlap the tail of kernel k (the
// last block's grid reduction and its cross-GPU round trip).  Every kernel launched through launch_k() therefore starts
// with pdl_enter(): wait for the predecessor, then allow the successor to be scheduled behind us.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }

bool pdl_enabled();   // engine.cu: lcgb200_set_pdl / LCGB200_PDL (default on)

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args&&... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1u : 0u;
	return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

struct OpBase {
	__device__ __forceinline__ bool active(const DevState*) const { return true; }
	__device__ __forceinline__ void begin(const DevState*) {}
	__device__ __forceinline__ void finish(DevState*, const double*) const {}
};

// ---- the generic fused vector kernel ----------------------------------------------------------------------
// Op interface:
//   static constexpr int NRED;      number of double reduction slots (0 = pure update, no grid reduction)
//   static constexpr int W;         elements per access on the main path (2 for real vectors, 1 for complex)
//   __device__ void begin(const DevState*);                       load scalars into the (kernel-local) op copy
//   __device__ bool active(const DevState*);                      false -> the whole kernel is a no-op (OpBase: true)
//   template<int V> __device__ void elem(size_t i, double* acc);  process V elements starting at i
//   __device__ void finish(DevState*, const double* tot);         scalar epilogue (one thread of the grid)
// PUSH (multi-GPU, NVLink transport): `push_src` is the vector this kernel produces and the next SpMV consumes; its
// boundary entries are pushed to the neighbours as they are written (T = element type of that vector).
template <class Op, bool PUSH = false, class T = double>
__global__ void __launch_bounds__(kThreads) k_vec(Op op_in, size_t n, DevState* st, double* partials, CommDev* comm = nullptr, const T* push_src = nullptr)
{
	pdl_enter();
	if (st_done(st)) return;
	Op op = op_in;
	if (!op.active(st)) return;
	op.begin(st);
	double acc[Op::NRED > 0 ? Op::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Op::NRED > 0 ? Op::NRED : 1); r++) acc[r] = 0.0;
	constexpr int W = Op::W;
	const size_t npack = n / W;
	const size_t stride = (size_t)gridDim.x * kThreads;
	unsigned long long hseq = 0; bool pushed = false;
	if (PUSH) hseq = comm->halo_seq + 1;
	for (size_t p = (size_t)blockIdx.x * kThreads + threadIdx.x; p < npack; p += stride)
	{
		op.template elem<W>(p * W, acc);
		if (PUSH) pushed |= push_elems<T, W>(comm, push_src, p * W, hseq);
	}
	if (W > 1)
	{
		const size_t tail = npack * W + (size_t)blockIdx.x * kThreads + threadIdx.x;
		if (tail < n)
		{
			op.template elem<1>(tail, acc);
			if (PUSH) pushed |= push_elems<T, 1>(comm, push_src, tail, hseq);
		}
	}
	if (PUSH) push_signal(comm, hseq, pushed);
	if (Op::NRED > 0)
	{
		double tot[Op::NRED > 0 ? Op::NRED : 1];
		if (grid_reduce<(Op::NRED > 0 ? Op::NRED : 1)>(acc, partials, &st->ticket, tot))
		{
			if (st->multi) { if (reduce_across_ranks(st, tot, Op::NRED)) op.finish(st, tot); }
			else if ((threadIdx.x & 31) == 0) op.finish(st, tot);
		}
	}
}

// ---- two dependent vector steps in one kernel ---------------------------------------------------------------------
// Every iteration ends with the pair "update (m, r, [z], reductions -> beta, loop head)" -> "direction (d = z + beta d)".  The
// second step needs ONE scalar of the first one's reduction, nothing else crosses threads: both steps walk the vectors
// with the same index mapping, so a thread re-reads in step 2 only what it wrote itself in step 1 (L2 hits at the sizes
// where this matters).  One cooperative launch (all blocks co-resident) runs both: step 1, single-pass grid reduction, the
// last block finishes the reduction (across GPUs too), runs the scalar epilogue and opens the gate the other blocks spin
// on; then step 2 (and the halo push).  Saves a kernel boundary per iteration — launch, ramp-up, tail — which is a fifth of
// the iteration on a 2 M-row system or an 8-GPU slab.
__device__ __forceinline__ int ld_acquire_gpu_s32(const int* p)
{
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_gpu_s32(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

template <class Op1, class Op2, bool PUSH, class T>
__global__ void __launch_bounds__(kThreads) k_vec2(Op1 op1_in, Op2 op2_in, size_t n, DevState* st, double* partials, CommDev* comm, const T* push_src)
{
	static_assert(Op1::W == Op2::W && Op1::NRED > 0, "same index mapping in both steps; the first one ends in a reduction");
	pdl_enter();
	if (st_done(st)) return;
	__shared__ int s_gate;
	if (threadIdx.x == 0) s_gate = ld_acquire_gpu_s32(&st->gate);
	constexpr int W = Op1::W;
	const size_t npack = n / W;
	const size_t stride = (size_t)gridDim.x * kThreads;
	const size_t first = (size_t)blockIdx.x * kThreads + threadIdx.x;
	{
		Op1 op = op1_in;
		op.begin(st);
		double acc[Op1::NRED];
#pragma unroll
		for (int r = 0; r < Op1::NRED; r++) acc[r] = 0.0;
		for (size_t p = first; p < npack; p += stride) op.template elem<W>(p * W, acc);
		if (W > 1) { const size_t tail = npack * W + first; if (tail < n) op.template elem<1>(tail, acc); }
		double tot[Op1::NRED];
		if (grid_reduce<Op1::NRED>(acc, partials, &st->ticket, tot))
		{	// first warp of the block that finished last
			if (st->multi) { if (reduce_across_ranks(st, tot, Op1::NRED)) op.finish(st, tot); }
			else if ((threadIdx.x & 31) == 0) op.finish(st, tot);
			__syncwarp();
			if ((threadIdx.x & 31) == 0) { __threadfence(); st_release_gpu_s32(&st->gate, s_gate + 1); }
		}
	}
	if (threadIdx.x == 0) { const int g0 = s_gate; while (ld_acquire_gpu_s32(&st->gate) == g0) __nanosleep(20); }
	__syncthreads();
	if (st_done(st)) return;   // the loop head of step 1 ended the solve: the direction step is skipped, as a separate launch would be
	{
		Op2 op = op2_in;
		op.begin(st);
		double none[1];
		unsigned long long hseq = 0; bool pushed = false;
		if (PUSH) hseq = comm->halo_seq + 1;
		for (size_t p = first; p < npack; p += stride)
		{
			op.template elem<W>(p * W, none);
			if (PUSH) pushed |= push_elems<T, W>(comm, push_src, p * W, hseq);
		}
		if (W > 1)
		{
			const size_t tail = npack * W + first;
			if (tail < n) { op.template elem<1>(tail, none); if (PUSH) pushed |= push_elems<T, 1>(comm, push_src, tail, hseq); }
		}
		if (PUSH) push_signal(comm, hseq, pushed);
	}
}

// multi-GPU: scalar epilogue after the totals in st->red have been all-reduced across ranks
template <class Op>
__global__ void k_finish(Op op_in, DevState* st)
{
	pdl_enter();
	if (st_done(st)) return;
	Op op = op_in;
	if (!op.active(st)) return;
	op.begin(st);
	double tot[kMaxRed];
	for (int r = 0; r < kMaxRed; r++) tot[r] = st->red[r];
	op.finish(st, tot);
}

inline int vec_grid(size_t n, int w)
{
	size_t packs = (n + w - 1) / w;
	size_t blocks = (packs + kThreads - 1) / kThreads;
	if (blocks < 1) blocks = 1;
	if (blocks > (size_t)kMaxBlocks) blocks = kMaxBlocks;
	return (int)blocks;
}

#define LCG_CUDA_CHECK(call)                                                                        \
	do {                                                                                            \
		cudaError_t e__ = (call);                                                                   \
		if (e__ != cudaSuccess) { lcgb200::set_error(#call, e__, __FILE__, __LINE__); throw lcgb200::CudaFailure(); } \
	} while (0)

struct CudaFailure {};
void set_error(const char* what, cudaError_t e, const char* file, int line);
void set_error_msg(const char* msg);

}  // namespace lcgb200
