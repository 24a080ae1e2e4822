// csr.cuh — built-in CSR operator: device layout, row tiles, and the streaming SpMV kernel with a fused
// per-row epilogue (dot products / vector updates) and the same single-pass grid reduction as k_vec.
//
// Layout in HBM: int32 row_ptr[n+1], int32 col[nnz (+pad)], T val[nnz (+pad)], base 0, rows in
// order — exactly the arrays the reference's samples hand to cusparseCreateCsr (sample8.cu:172-173) — plus one
// int4 {row_begin, row_end, nnz_begin_aligned, nnz_end} per row tile.  A tile is a run of consecutive rows
// whose non-zeros (from the 4-aligned start) fit one shared-memory stage; the tile builder caps the row count
// at a multiple of the row groups of a block so that the compute phase has no ragged last pass.
//
// Kernel (sm_100a): a persistent grid (4 CTAs x 148 SMs, 56 registers/thread), warp-specialised.  The constants below
// are the best of the sweeps in profiles/sweep_r01.txt (more resident consumer warps beat deeper rings).
//   producer  one elected thread of an extra warp streams the tile's col / val / row_ptr slices HBM -> shared
//             memory with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx, L2 evict-first hint: the
//             matrix is read exactly once per SpMV and must not push the vectors out of L2) into a ring of
//             kStages stages, running up to kStages-1 tiles ahead of the consumers;
//   consumers 8 warps wait on the stage's "full" mbarrier, LPR lanes per row (1..32, picked from the mean row
//             length) walk their row out of shared memory and gather x[col] through L1/L2 (lanes of a warp sit
//             on consecutive rows, so for banded/stencil matrices a warp-wide gather touches 2-3 lines), then
//             release the stage through its "empty" mbarrier.  No block-wide barrier in the steady state.
//   epilogue  the row result feeds the fused reductions (p.Ap, r0~.Ap, As.s, As.As, ...) without re-reading y.
// Tiles are dealt to CTAs in chunks of consecutive tiles: inside a chunk the gathered x of neighbouring rows is
// re-used out of L1, and at any moment all CTAs work inside one band of rows, so the band's x stays L2-resident.
#pragma once
#include "common.cuh"
#include <type_traits>

namespace lcgb200 {

#ifndef LCG_TILE_NNZ
#define LCG_TILE_NNZ 1792
#endif
#ifndef LCG_STAGES
#define LCG_STAGES 2
#endif
#ifndef LCG_CTAS
#define LCG_CTAS 4
#endif
#ifndef LCG_UNROLL
#define LCG_UNROLL 8
#endif
constexpr int kTileNnzReal = LCG_TILE_NNZ;   // staged non-zeros per tile (14 KB val + 7 KB col): 256 rows x 7 or 64 rows x 27-28
constexpr int kTileNnzCplx = 1024;   // 16 KB val + 4 KB col
constexpr int kTileRows = 512;       // max rows per tile (bounds the row_ptr slice in shared memory)
constexpr int kStages = LCG_STAGES;           // shared-memory ring depth; 4 CTAs x 2 stages x 22.5 KB leave ~45 KB of L1 per SM for the gathered x
constexpr int kGatherUnroll = LCG_UNROLL;     // row entries per lane whose loads are issued back to back
constexpr int kSpmvThreads = kThreads + 32;   // 8 consumer warps + 1 producer warp
constexpr int kSpmvCtasPerSm = LCG_CTAS;

struct PatMarch;   // plan of the plane-marching row-pattern kernel (below)
struct PatBox;     // plan of the box row-pattern kernel (below)
template <class T>
struct CsrDev {
	int n_rows = 0, n_cols = 0, nnz = 0, n_tiles = 0, lpr = 1, chunk = 1;
	// row block of a partitioned system on the NVLink transport: tiles [n_interior, n_tiles) reference ghost columns
	// (>= n_rows) and are read from the halo mailbox once the neighbours' pushes have landed; -1 = not in use
	int n_interior = -1;
	CommDev* comm = nullptr;   // the transport state of THIS block's halo plan (a partitioned transpose has its own)
	const int* row_ptr = nullptr;
	const int* col = nullptr;
	const T* val = nullptr;
	const int4* tiles = nullptr;
	// dictionary-compressed copy (real operators, optional): code = value index | offset index << 8
	const unsigned short* code = nullptr;
	const double* vdict = nullptr; const int* odict = nullptr;
	const int4* dtiles = nullptr; int n_dtiles = 0, dchunk = 1, dlpr = 1;
	// row-pattern copy (real operators, optional): one pattern id per ROW + a table of the distinct rows
	// (csr.cuh "row-pattern operator": chains of offsets S apart, one byte per warp work item)
	// pat_chain: chains | info of all patterns, one device array;
	// pat_bitem / pat_segs / pat_march: block items, march segments and plan of k_spmv_pat_march (host pointer, null = no plan)
	// pat_thread: a byte per thread of a warp item (the pattern its R rows share, 255 = mixed)
	const unsigned char* pat = nullptr; const unsigned char* pat_thread = nullptr; const void* pat_chain = nullptr;
	const unsigned char* pat_bitem = nullptr; const void* pat_segs = nullptr; const PatMarch* pat_march = nullptr;
	// pat_box / pat_box_flags: plan (host pointer, null = no plan) and one byte per thread item of k_spmv_pat_box
	const PatBox* pat_box = nullptr; const unsigned char* pat_box_flags = nullptr;
	int n_pat = 0, pat_maxch = 0, pat_stride = 0, pat_nib = 0, pat_items = 0;
};

template <class T> struct TileCfg;
template <> struct TileCfg<double> { static constexpr int NNZ = kTileNnzReal; static constexpr int CTAS = kSpmvCtasPerSm; };
// complex rows carry twice the registers per entry: 2 CTAs per SM (112 registers) keep the gather loop spill-free
template <> struct TileCfg<ZF> { static constexpr int NNZ = kTileNnzReal; static constexpr int CTAS = kSpmvCtasPerSm > 2 ? 2 : kSpmvCtasPerSm; };   // 8-byte entries like double, complex arithmetic
template <> struct TileCfg<double2> { static constexpr int NNZ = kTileNnzCplx; static constexpr int CTAS = kSpmvCtasPerSm > 2 ? 2 : kSpmvCtasPerSm; };

// shared-memory stage layout (bytes): val | col | row_ptr slice
template <class T> struct StageCfg {
	static constexpr int NNZ = TileCfg<T>::NNZ;
	static constexpr int VAL_OFF = 0;
	static constexpr int COL_OFF = NNZ * (int)sizeof(T);
	static constexpr int ROW_OFF = COL_OFF + NNZ * 4;
	static constexpr int BYTES = (ROW_OFF + (kTileRows + 8) * 4 + 127) & ~127;
	static constexpr int TOTAL = BYTES * kStages;
};

__device__ __forceinline__ double mulacc(double acc, double a, double x) { return fma(a, x, acc); }
#ifdef LCG_REFORDER
// reference order: the complex product is formed first (re = ac - bd, im = ad + bc), then added to the running sum
__device__ __forceinline__ double2 mulacc(double2 acc, double2 a, double2 x) { return zadd(acc, zmul(a, x)); }
#else
__device__ __forceinline__ double2 mulacc(double2 acc, double2 a, double2 x)
{
	acc.x = fma(a.x, x.x, acc.x); acc.x = fma(-a.y, x.y, acc.x);
	acc.y = fma(a.x, x.y, acc.y); acc.y = fma(a.y, x.x, acc.y);
	return acc;
}
#endif
// single-precision complex entries: products and row sums in float (as cusparseSpMV does for CUDA_C_32F)
__device__ __forceinline__ ZF mulacc(ZF acc, ZF a, ZF x)
{
	acc.x = fmaf(a.x, x.x, acc.x); acc.x = fmaf(-a.y, x.y, acc.x);
	acc.y = fmaf(a.x, x.y, acc.y); acc.y = fmaf(a.y, x.x, acc.y);
	return acc;
}
__device__ __forceinline__ ZF tzero(ZF) { return ZF(0.f, 0.f); }
__device__ __forceinline__ ZF tconj(ZF a) { return ZF(a.x, -a.y); }
__device__ __forceinline__ ZF tadd(ZF a, ZF b) { return ZF(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ ZF tshfl_xor(ZF v, int o) { return ZF(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o)); }
__device__ __forceinline__ ZF tldg(const ZF* p) { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); return ZF(t.x, t.y); }
__device__ __forceinline__ double tzero(double) { return 0.0; }
__device__ __forceinline__ double2 tzero(double2) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double tconj(double a) { return a; }
__device__ __forceinline__ double2 tconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double tadd(double a, double b) { return a + b; }
__device__ __forceinline__ double2 tadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double tshfl_xor(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double2 tshfl_xor(double2 v, int o)
{
	return make_double2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o));
}
__device__ __forceinline__ double tldg(const double* p) { return __ldg(p); }
__device__ __forceinline__ double2 tldg(const double2* p) { return __ldg(p); }

// ---- mbarrier / TMA bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	uint32_t ok;
	do
	{
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	} while (!ok);
}
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
// global -> shared bulk copy, completion counted in bytes on `bar`; dst/src 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
		:: "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
// consumer-only barrier (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kThreads) : "memory"); }

int spmv_grid_limit(int ctas_per_sm);   // resident CTAs of k_spmv on the current device (SMs x CTAs per SM), engine.cu

// cudaFuncSetAttribute applies to ONE device: remember per (kernel instantiation, device) whether the dynamic shared-memory
// limit has been raised (a process may drive several GPUs)
struct PerDeviceOnce {
	bool done[64] = {false};
	bool first()
	{
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
		if (done[dev]) return false;
		done[dev] = true;
		return true;
	}
};

// One staged tile: LPR lanes per row walk the rows out of shared memory and gather x.  GHOST: columns >= n_local are
// ghost entries and are read in place from the halo mailbox (`ghost` is biased by -n_local, so it is indexed by the column).
template <class T, int LPR, bool CONJ, bool GHOST, class Epi>
__device__ __forceinline__ void spmv_tile_rows(const T* __restrict__ sval, const int* __restrict__ scol, const int* __restrict__ srow,
	int r0, int nrows, int k0, const T* __restrict__ x, const T* ghost, int n_local, T* __restrict__ y, Epi& epi, double* acc, int group, int lane)
{
	constexpr int NG = kThreads / LPR;          // row groups per block
	for (int rb = 0; rb < nrows; rb += NG)
	{
		const int r = rb + group;
		int kb = 0, ke = 0;
		if (r < nrows) { kb = srow[r] - k0; ke = srow[r + 1] - k0; }
		T sum = tzero(T());
		// kGatherUnroll entries per lane at a time, all their loads issued before the first use: the row walk
		// is bound by the latency of the gathered x (L2 for most stencil neighbours), not by issue slots
		for (int j0 = kb + lane; j0 < ke; j0 += LPR * kGatherUnroll)
		{
			int cidx[kGatherUnroll]; T a[kGatherUnroll], xv[kGatherUnroll];
#pragma unroll
			for (int u = 0; u < kGatherUnroll; u++)
			{
				const int j = j0 + u * LPR;
				const bool ok = j < ke;
				cidx[u] = ok ? scol[j] : -1;
				a[u] = ok ? sval[j] : tzero(T());
			}
#pragma unroll
			for (int u = 0; u < kGatherUnroll; u++)
			{
				if (GHOST) xv[u] = cidx[u] >= 0 ? tldg((cidx[u] >= n_local ? ghost : x) + cidx[u]) : tzero(T());
				else xv[u] = cidx[u] >= 0 ? tldg(x + cidx[u]) : tzero(T());
			}
#pragma unroll
			for (int u = 0; u < kGatherUnroll; u++)
			{
				if (CONJ) a[u] = tconj(a[u]);
				sum = mulacc(sum, a[u], xv[u]);
			}
		}
#pragma unroll
		for (int o = LPR / 2; o > 0; o >>= 1) sum = tadd(sum, tshfl_xor(sum, o));
		if (lane == 0 && r < nrows)
		{
			y[r0 + r] = sum;
			epi.row(r0 + r, sum, tldg(x + r0 + r), acc);
		}
	}
}

// Consumer side of k_spmv: chunks c, c + gridDim.x, ... below c_end.  `idx` counts the staged tiles this CTA has consumed
// (the ring position shared with the producer), so that two calls can split the chunk range between them.
template <class T, int LPR, bool CONJ, bool GHOST, class Epi>
__device__ __forceinline__ void spmv_consume(const CsrDev<T>& A, const T* __restrict__ x, const T* ghost, T* __restrict__ y, Epi& epi, double* acc,
	unsigned char* smem, unsigned long long* s_bar, T* s_long, int& idx, int& c, const int c_end, const int tid)
{
	typedef StageCfg<T> SC;
	constexpr int TN = SC::NNZ;
	const int group = tid / LPR, lane = tid % LPR;
	for (; c < c_end; c += gridDim.x)
	{
		const int t1 = min((c + 1) * A.chunk, A.n_tiles);
		for (int tile = c * A.chunk; tile < t1; tile++)
		{
			const int4 td = __ldg(A.tiles + tile);
			const int r0 = td.x, nrows = td.y - td.x, k0 = td.z, k1 = td.w;
			if (k1 - k0 > TN)
			{	// one row longer than a stage: stream it straight from global memory
				T part = tzero(T());
				for (int k = A.row_ptr[r0] + tid; k < k1; k += kThreads)
				{
					T a = A.val[k]; if (CONJ) a = tconj(a);
					const int cc = A.col[k];
					part = mulacc(part, a, tldg((GHOST && cc >= A.n_rows ? ghost : x) + cc));
				}
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) part = tadd(part, tshfl_xor(part, o));
				if ((tid & 31) == 0) s_long[tid >> 5] = part;
				consumer_sync();
				if (tid == 0)
				{
					T tot = s_long[0];
					for (int w = 1; w < kThreads / 32; w++) tot = tadd(tot, s_long[w]);
					y[r0] = tot;
					epi.row(r0, tot, tldg(x + r0), acc);
				}
				consumer_sync();
				continue;
			}
			const int s = idx % kStages;
			mbar_wait(smem_u32(&s_bar[s]), (uint32_t)((idx / kStages) & 1));
			const unsigned char* base = smem + (size_t)s * SC::BYTES;
			const T* sval = reinterpret_cast<const T*>(base + SC::VAL_OFF);
			const int* scol = reinterpret_cast<const int*>(base + SC::COL_OFF);
			const int* srow = reinterpret_cast<const int*>(base + SC::ROW_OFF) + (r0 & 3);
			spmv_tile_rows<T, LPR, CONJ, GHOST, Epi>(sval, scol, srow, r0, nrows, k0, x, ghost, A.n_rows, y, epi, acc, group, lane);
			__syncwarp();
			if ((tid & 31) == 0) mbar_arrive(smem_u32(&s_bar[kStages + s]));
			idx++;
		}
	}
}

// Epi interface:
//   static constexpr int NRED;
//   __device__ void begin(const DevState*);
//   __device__ void row(int i, T yi, T xi, double* acc);   called once per row by one lane (may write vectors)
//   __device__ void finish(DevState*, const double* tot);
// PART: row block of a partitioned system on the NVLink transport (receive half of the halo exchange in the boundary
// tiles); a separate instantiation, so that the single-GPU kernel carries none of it.
template <class T, int LPR, bool CONJ, class Epi, bool PART = false>
__global__ void __launch_bounds__(kSpmvThreads, TileCfg<T>::CTAS) k_spmv(CsrDev<T> A, const T* __restrict__ x, T* __restrict__ y, Epi epi_in,
	DevState* st, double* partials)
{
	// Programmatic dependent launch: this grid may be scheduled while its predecessor (the kernel that wrote x and the
	// iteration scalars) is still in its reduction tail.  Nothing the predecessor writes is touched before pdl_wait(): the
	// barriers are set up and the producer already streams the first stages of the MATRIX (which no kernel of the solve
	// writes) into shared memory, so the consumers find their first tiles waiting when the predecessor completes.
	pdl_trigger();
	typedef StageCfg<T> SC;
	constexpr int TN = SC::NNZ;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) unsigned long long s_bar[2 * kStages];   // [0,S) full, [S,2S) empty
	__shared__ T s_long[kThreads / 32];

	const int tid = threadIdx.x;
	const bool producer = tid >= kThreads;
	if (tid == 0)
	{
		for (int s = 0; s < kStages; s++) { mbar_init(smem_u32(&s_bar[s]), 1); mbar_init(smem_u32(&s_bar[kStages + s]), kThreads / 32); }
		mbar_fence_init();
	}
	__syncthreads();

	Epi epi = epi_in;
	double acc[Epi::NRED > 0 ? Epi::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Epi::NRED > 0 ? Epi::NRED : 1); r++) acc[r] = 0.0;

	const int n_chunks = (A.n_tiles + A.chunk - 1) / A.chunk;

	if (producer)
	{
		if (tid != kThreads) { pdl_wait(); if (st_done(st)) return; }
		else
		{
			const uint64_t pol = l2_evict_first_policy();
			int idx = 0;
			bool waited = false, stop = false;
			for (int c = blockIdx.x; c < n_chunks && !stop; c += gridDim.x)
			{
				const int t1 = min((c + 1) * A.chunk, A.n_tiles);
				for (int tile = c * A.chunk; tile < t1; tile++)
				{
					const int4 td = __ldg(A.tiles + tile);
					if (td.w - td.z > TN) continue;   // over-long row: the consumers stream it from global memory
					const int s = idx % kStages, j = idx / kStages;
					if (j > 0 && !waited)
					{	// the first kStages tiles went out ahead of the predecessor's completion; from here on the consumers are needed
						pdl_wait(); waited = true;
						if (st_done(st)) { stop = true; break; }
					}
					if (j > 0) mbar_wait(smem_u32(&s_bar[kStages + s]), (uint32_t)((j - 1) & 1));
					const uint32_t cnt = (uint32_t)((td.w - td.z + 3) & ~3);
					const int ra = td.x & ~3;
					const uint32_t rcnt = (uint32_t)((td.y - ra + 1 + 3) & ~3);
					const uint32_t full = smem_u32(&s_bar[s]);
					const uint32_t base = smem_u32(smem + (size_t)s * SC::BYTES);
					mbar_expect_tx(full, cnt * (uint32_t)(sizeof(T) + 4) + rcnt * 4u);
					if (cnt > 0)
					{	// a tile made only of empty rows has nothing to stream but its row_ptr slice
						bulk_g2s(base + SC::VAL_OFF, A.val + td.z, cnt * (uint32_t)sizeof(T), full, pol);
						bulk_g2s(base + SC::COL_OFF, A.col + td.z, cnt * 4u, full, pol);
					}
					bulk_g2s(base + SC::ROW_OFF, A.row_ptr + ra, rcnt * 4u, full, pol);
					idx++;
				}
			}
			if (!waited) { pdl_wait(); if (st_done(st)) stop = true; }
			if (stop)
			{	// the solve is over and the consumers have left: let the copies already in flight land before this CTA's shared memory is released
				for (int s = 0; s < kStages && s < idx; s++) mbar_wait(smem_u32(&s_bar[s]), 0u);
				return;
			}
		}
	}
	else
	{
		pdl_wait();
		if (st_done(st)) return;
		epi.begin(st);
		int idx = 0, c = blockIdx.x;
		if (PART && A.n_interior >= 0)
		{	// partitioned row block (NVLink transport): the chunks made of interior tiles only run the very loop of the single-GPU
			// kernel; then this warp waits until every neighbour's push for this exchange has landed in my mailbox and walks the
			// remaining chunks with ghost columns read in place from the mailbox
			spmv_consume<T, LPR, CONJ, false, Epi>(A, x, nullptr, y, epi, acc, smem, s_bar, s_long, idx, c, A.n_interior / A.chunk, tid);
			CommDev* cd = A.comm;
			const unsigned long long hseq = cd->halo_seq;
			bool ok = true;
			if ((tid & 31) < cd->n_peers && cd->recv_count[tid & 31] > 0)
				ok = spin_until(&cd->win[cd->rank]->halo_flag[cd->peer_rank[tid & 31]], hseq, st->spin_timeout_ns);
			ok = __all_sync(0xffffffffu, ok);
			if (!ok && (tid & 31) == 0) { st->ret = RC_UNKNOWN; st->done = 1; cd->abort_flag = 1; }
			const T* ghost = mailbox_of<T>(cd->win[cd->rank], hseq, cd->n_ghost) - cd->n_local;
			spmv_consume<T, LPR, CONJ, true, Epi>(A, x, ghost, y, epi, acc, smem, s_bar, s_long, idx, c, n_chunks, tid);
		}
		else spmv_consume<T, LPR, CONJ, false, Epi>(A, x, nullptr, y, epi, acc, smem, s_bar, s_long, idx, c, n_chunks, tid);
	}
	if (PART && A.n_interior >= 0)
	{	// the block that leaves last tells every sender that this exchange's mailbox buffer has been consumed
		__syncthreads();
		if (tid == 0)
		{
			CommDev* cd = A.comm;
			if (atomicAdd(&cd->ticket2, 1u) == gridDim.x - 1)
			{
				const unsigned long long hseq = cd->halo_seq;
				for (int p = 0; p < cd->n_peers; p++)
					if (cd->recv_count[p] > 0) st_relaxed_sys(&cd->win[cd->peer_rank[p]]->halo_ack[cd->rank], hseq);
				cd->ticket2 = 0u;
			}
		}
	}
	if (Epi::NRED > 0)
	{
		double tot[Epi::NRED > 0 ? Epi::NRED : 1];
		if (grid_reduce<(Epi::NRED > 0 ? Epi::NRED : 1)>(acc, partials, &st->ticket, tot))
		{
			if (st->multi) { if (reduce_across_ranks(st, tot, Epi::NRED)) epi.finish(st, tot); }
			else if ((threadIdx.x & 31) == 0) epi.finish(st, tot);
		}
	}
}

// ---- dictionary-compressed operator ------------------------------------------------------------------------------
// Matrices with few distinct values and few distinct (col - row) offsets — constant-coefficient stencils, and their
// row blocks after the ghost remap — are stored a second time as ONE 16-bit code per entry (value index | offset index
// << 8) plus two dictionaries of <= 256 entries: 2 bytes per non-zero instead of 12 stream from HBM.  The kernel is
// the same producer/consumer pipeline; consumers decode through the dictionaries held in shared memory.  Entries,
// per-lane accumulation order and butterfly are those of k_spmv, so y and the fused dots are bitwise identical.
constexpr int kDictTileNnz = 7168;    // 14 KB of codes per stage: 256 rows x 27-28 or 512 rows x 7
struct DictStage {
	static constexpr int CODE_OFF = 0;
	static constexpr int ROW_OFF = kDictTileNnz * 2;
	static constexpr int BYTES = (ROW_OFF + (kTileRows + 8) * 4 + 127) & ~127;
	static constexpr int TOTAL = BYTES * kStages;
};

template <int LPR, class Epi>
__global__ void __launch_bounds__(kSpmvThreads, kSpmvCtasPerSm) k_spmv_dict(CsrDev<double> A, const double* __restrict__ x, double* __restrict__ y, Epi epi_in,
	DevState* st, double* partials)
{
	pdl_enter();
	if (st_done(st)) return;
	typedef DictStage SC;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) unsigned long long s_bar[2 * kStages];
	__shared__ double s_vdict[256];
	__shared__ int s_odict[256];

	const int tid = threadIdx.x;
	const bool producer = tid >= kThreads;
	if (tid < 256) { s_vdict[tid] = A.vdict[tid]; s_odict[tid] = A.odict[tid]; }
	if (tid == 0)
	{
		for (int s = 0; s < kStages; s++) { mbar_init(smem_u32(&s_bar[s]), 1); mbar_init(smem_u32(&s_bar[kStages + s]), kThreads / 32); }
		mbar_fence_init();
	}
	__syncthreads();

	Epi epi = epi_in;
	double acc[Epi::NRED > 0 ? Epi::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Epi::NRED > 0 ? Epi::NRED : 1); r++) acc[r] = 0.0;
	const int n_chunks = (A.n_dtiles + A.dchunk - 1) / A.dchunk;

	if (producer)
	{
		if (tid == kThreads)
		{
			const uint64_t pol = l2_evict_first_policy();
			int idx = 0;
			for (int c = blockIdx.x; c < n_chunks; c += gridDim.x)
			{
				const int t1 = min((c + 1) * A.dchunk, A.n_dtiles);
				for (int tile = c * A.dchunk; tile < t1; tile++)
				{
					const int4 td = __ldg(A.dtiles + tile);
					const int s = idx % kStages, j = idx / kStages;
					if (j > 0) mbar_wait(smem_u32(&s_bar[kStages + s]), (uint32_t)((j - 1) & 1));
					const uint32_t cnt = (uint32_t)((td.w - td.z + 7) & ~7);
					const int ra = td.x & ~3;
					const uint32_t rcnt = (uint32_t)((td.y - ra + 1 + 3) & ~3);
					const uint32_t full = smem_u32(&s_bar[s]);
					const uint32_t base = smem_u32(smem + (size_t)s * SC::BYTES);
					mbar_expect_tx(full, cnt * 2u + rcnt * 4u);
					if (cnt > 0) bulk_g2s(base + SC::CODE_OFF, A.code + td.z, cnt * 2u, full, pol);
					bulk_g2s(base + SC::ROW_OFF, A.row_ptr + ra, rcnt * 4u, full, pol);
					idx++;
				}
			}
		}
	}
	else
	{
		epi.begin(st);
		constexpr int NG = kThreads / LPR;
		const int group = tid / LPR, lane = tid % LPR;
		int idx = 0;
		for (int c = blockIdx.x; c < n_chunks; c += gridDim.x)
		{
			const int t1 = min((c + 1) * A.dchunk, A.n_dtiles);
			for (int tile = c * A.dchunk; tile < t1; tile++)
			{
				const int4 td = __ldg(A.dtiles + tile);
				const int r0 = td.x, nrows = td.y - td.x, k0 = td.z;
				const int s = idx % kStages;
				mbar_wait(smem_u32(&s_bar[s]), (uint32_t)((idx / kStages) & 1));
				const unsigned char* base = smem + (size_t)s * SC::BYTES;
				const unsigned short* scode = reinterpret_cast<const unsigned short*>(base + SC::CODE_OFF);
				const int* srow = reinterpret_cast<const int*>(base + SC::ROW_OFF) + (r0 & 3);
				for (int rb = 0; rb < nrows; rb += NG)
				{
					const int r = rb + group;
					const int row = r0 + r;
					int kb = 0, ke = 0;
					if (r < nrows) { kb = srow[r] - k0; ke = srow[r + 1] - k0; }
					double sum = 0.0;
					for (int j0 = kb + lane; j0 < ke; j0 += LPR * kGatherUnroll)
					{
						int cidx[kGatherUnroll]; double a[kGatherUnroll], xv[kGatherUnroll];
#pragma unroll
						for (int u = 0; u < kGatherUnroll; u++)
						{
							const int j = j0 + u * LPR;
							const bool ok = j < ke;
							const unsigned int cd = ok ? (unsigned int)scode[j] : 0u;
							cidx[u] = ok ? row + s_odict[cd >> 8] : -1;
							a[u] = ok ? s_vdict[cd & 255u] : 0.0;
						}
#pragma unroll
						for (int u = 0; u < kGatherUnroll; u++) xv[u] = cidx[u] >= 0 ? tldg(x + cidx[u]) : 0.0;
#pragma unroll
						for (int u = 0; u < kGatherUnroll; u++) sum = mulacc(sum, a[u], xv[u]);
					}
#pragma unroll
					for (int o = LPR / 2; o > 0; o >>= 1) sum = tadd(sum, tshfl_xor(sum, o));
					if (lane == 0 && r < nrows)
					{
						y[row] = sum;
						epi.row(row, sum, tldg(x + row), acc);
					}
				}
				__syncwarp();
				if ((tid & 31) == 0) mbar_arrive(smem_u32(&s_bar[kStages + s]));
				idx++;
			}
		}
	}
	if (Epi::NRED > 0)
	{
		double tot[Epi::NRED > 0 ? Epi::NRED : 1];
		if (grid_reduce<(Epi::NRED > 0 ? Epi::NRED : 1)>(acc, partials, &st->ticket, tot))
		{
			if (st->multi) { if (reduce_across_ranks(st, tot, Epi::NRED)) epi.finish(st, tot); }
			else if ((threadIdx.x & 31) == 0) epi.finish(st, tot);
		}
	}
}

// ---- row-pattern operator ----------------------------------------------------------------------------------------
// Constant-coefficient discretisations have only a handful of DISTINCT ROWS once a row is written as its sequence of
// (col - row, value) pairs: interior, faces, edges, corners (27 for a 3-D stencil; a few more after the ghost remap of a
// row block).  Such a matrix is stored a third time as ONE BYTE PER ROW (its pattern id) plus the table of patterns:
// the SpMV streams x, y and n bytes of ids — the 12 bytes per non-zero of CSR disappear, and what bounds the kernel is
// how the gathers of x are served (round 1: 27 per row, LSU data pipe 84 % busy, 0.2 of HBM).
//
// 1. Gathers shared between rows (both kernels).  The entries of a pattern are grouped into CHAINS: up to kPatChainLen
//    offsets in arithmetic progression with the matrix-wide stride S (for a 3-D stencil S = nx: the entries (dx, dz)
//    fixed, dy = -1, 0, +1).  A thread owns R rows that are S apart (row, row + S, ..., row + (R-1) S: a column of R grid
//    points in y), so the x values one chain needs for all R rows are R + m - 1 reads instead of R * m — x[row + off +
//    u S] serves row q = u - t through chain entry t.  27-point stencil, R = 8: 9 chains x 10 reads for 8 rows = 11.25
//    per row instead of 27, and one 32-byte table read per chain instead of one per entry.  Lanes of a warp sit on 32
//    consecutive rows.  Work item of a warp = R x 32 rows: rows (A R + q) S + 32 ib + lane.
// 2. Rows that differ.  A byte per THREAD of an item (pat_thread, loaded one round ahead) holds the pattern its R rows share
//    — the rule inside a stencil; lanes at the ends of a grid line simply hold another id.  255 = the rows differ: patterns
//    that are SUBSETS of one longer pattern (a column of grid points that starts on a face: the face row is the interior row
//    minus the entries that leave the grid; pat_host.h: pat_build_masks) take the longer pattern's chains with a 64-bit
//    presence mask per row, reads still shared; anything else goes row by row through the same chain table.
// 3. k_spmv_pat (any matrix with <= 253 distinct rows): the reads are plain loads through L1.  The 8 warps of
//    a block take the items of a round in rotation, so the slower items (ends of grid lines) do not pile up on two warps.
//    Measured at 27-point 256^3 (profiles/README_r02.md): 0.122 ms (round 1: 0.222), LSU data pipe 56 %, issue slots 56 %
//    busy, 1015 instructions per 256 rows — bound by instruction issue and load latency, not by bytes (DRAM: 0.24 GB).
// 4. k_spmv_pat_march (opt-in, LCGB200_PAT_MARCH=1; stride a multiple of 128, chain offsets on planes S2 apart — 3-D grids
//    with nx % 128 == 0 and ny % 8 == 0): a thread block of 8 consumer warps + 1 producer warp MARCHES along S2 (z).  The
//    producer's lanes copy the lines of one plane window — the x values all 8 warps need from one plane: (wy R + 2) lines x
//    (32 wx + 8) values — into a ring of shared-memory stages with TMA bulk copies (full/empty mbarriers, as k_spmv does
//    for the matrix).  The window of plane p of item k is the window of plane p - 1 of item k + 1, so ONE window is loaded
//    per item instead of G = 3: x crosses L2 -> SM 1.3 times per SpMV instead of 4 (measured 222 MB instead of 545 MB),
//    and the consumers read x out of shared memory with immediate offsets.  Measured 0.175 ms: the generic chain walk costs
//    as many instructions as in k_spmv_pat and a block advances at the pace of its slowest warp, so on one B200 — where x
//    stays in L2 anyway — the plain-load kernel wins; kept for systems whose x does not fit L2.
// 5. k_spmv_pat_box (default where it applies: the geometry pattern is a dense box — a 27-point stencil — and the grid lines and
//    planes align with the threads' columns, i.e. ny % 8 == 0): see the comment at the kernel.  0.083 ms at 27-point 256^3.
// Entries are accumulated chain by chain (box: plane by plane, line by line) with fma — a different order from the CSR row
// order: y agrees with the plain copy to rounding, not bitwise.
constexpr int kPatRows = 8;            // R: rows (S apart) a thread computes together
constexpr int kPatChainLen = 3;        // entries per chain (v[3])
constexpr int kPatMaxChains = 1536;    // chain table entries held in shared memory (48 KB)
constexpr int kPatDefaultStride = 256; // S when no pair of offsets repeats (all chains have one entry)
constexpr int kPatSpan = 8;            // extra values per window line of the marching kernel
constexpr int kPatMaxPlanes = 4;
constexpr int kPatMaxStages = 8;       // window stages of the marching kernel: planes + windows loaded ahead
// m = entries | line shift << 4 | position in the window line << 8 | plane << 16 (pat_host.h: PatChainH)
struct __align__(16) PatChain { double v[kPatChainLen]; int off; int m; };   // 32 bytes: two 128-bit shared-memory reads
static_assert(sizeof(PatChain) == 32, "PatChain is read as two 16-byte words");
// per pattern: presence mask over the chains of pattern `sup` (bit 3 c + t), chains | (index of offset 0 in the LAST chain, +1) << 8
struct __align__(16) PatInfo { unsigned long long mask; int info; int sup; };
// the marching kernel's plan (pat_host.h: PatMarchH) as the kernel gets it
struct PatMarch {
	int gpat = -1;      // geometry pattern, -1 = no plan
	int G = 0, S2 = 0, o0 = 0, nlines = 0, wx = 0, wy = 0, dAb = 0, n_segs = 0;
	int nst = 0;            // window stages in shared memory: G + the windows loaded ahead
	unsigned int gbegin = 0;   // chains of group g: [byte g, byte g + 1) of gbegin | gend << 32 ... (8 bits each, G <= 4: 5 bytes)
	unsigned int gend4 = 0;    // the fifth byte (end of the last group)
	unsigned int gplane = 0;   // plane of group g: bits [2 g, 2 g + 2)
	__host__ __device__ int group_begin(int g) const { return g < 4 ? (int)((gbegin >> (8 * g)) & 255u) : (int)gend4; }
	__host__ __device__ int group_plane(int g) const { return (int)((gplane >> (2 * g)) & 3u); }
};

__device__ __forceinline__ void pat_chain_load(const PatChain* c, double& v0, double& v1, double& v2, int& off, int& m)
{
	const double2 a = reinterpret_cast<const double2*>(c)[0];
	const double2 b = reinterpret_cast<const double2*>(c)[1];
	v0 = a.x; v1 = a.y; v2 = b.x;
	const long long om = __double_as_longlong(b.y);
	off = (int)(om & 0xffffffffll); m = (int)(om >> 32);
}

// The R rows of one thread through plain loads (paragraphs 1-3).  tcode = the pattern all R rows share, 255 = look at
// the rows.  Lanes with lane_on == false do nothing.
template <class Epi>
__device__ __forceinline__ void pat_item_ldg(const unsigned char* __restrict__ pat, const double* __restrict__ x, double* __restrict__ y,
	const PatChain* s_ch, const PatInfo* s_info, int maxch, int S, int n_rows, int row0, int tcode, bool lane_on, Epi& epi, double* acc)
{
	constexpr int R = kPatRows;
	if (!lane_on) return;
	int p = tcode;
	bool chained = true, masked = false;
	unsigned long long mk[R];
	if (tcode == 255)
	{
		int p0 = -1, sup = -1;
		masked = true;
#pragma unroll
		for (int q = 0; q < R; q++)
		{
			const int row = row0 + q * S;
			const int pq = row < n_rows ? (int)pat[row] : -1;
			if (q == 0) p0 = pq;
			chained = chained && pq >= 0 && pq == p0;
			mk[q] = 0ull;
			if (pq >= 0)
			{	// rows beyond the matrix keep an empty mask: no loads, no store
				const PatInfo pi = s_info[pq];
				if (sup < 0) sup = pi.sup;
				masked = masked && pi.sup == sup && pi.mask != 0ull;
				mk[q] = pi.mask;
			}
		}
		p = chained ? p0 : sup;
	}
	if (chained)
	{
		const int info = s_info[p].info;
		const int nch = info & 255, t0 = (info >> 8) - 1;
		const PatChain* ch = s_ch + p * maxch;
		double sum[R];
#pragma unroll
		for (int q = 0; q < R; q++) sum[q] = 0.0;
		for (int c = 0; c < nch; c++)
		{
			double v0, v1, v2; int off, m;
			pat_chain_load(ch + c, v0, v1, v2, off, m);
			m &= 3;
			const int xb = row0 + off;   // 32-bit element indices (try_patterns checks the range): one add + one widening multiply-add per load
			double xl[R + 2];
#pragma unroll
			for (int u = 0; u < R; u++) xl[u] = __ldg(x + (xb + u * S));
			xl[R] = m > 1 ? __ldg(x + (xb + R * S)) : 0.0;
			xl[R + 1] = m > 2 ? __ldg(x + (xb + (R + 1) * S)) : 0.0;
#pragma unroll
			for (int q = 0; q < R; q++) sum[q] = fma(v0, xl[q], sum[q]);
			if (m > 1)
			{
#pragma unroll
				for (int q = 0; q < R; q++) sum[q] = fma(v1, xl[q + 1], sum[q]);
			}
			if (m > 2)
			{
#pragma unroll
				for (int q = 0; q < R; q++) sum[q] = fma(v2, xl[q + 2], sum[q]);
			}
			if (c == nch - 1 && t0 >= 0)
			{	// the last chain holds the diagonal when the row has one: its values double as x[row] for the epilogue
#pragma unroll
				for (int q = 0; q < R; q++)
				{
					const int row = row0 + q * S;
					y[row] = sum[q];
					epi.row(row, sum[q], t0 == 0 ? xl[q] : (t0 == 1 ? xl[q + 1] : xl[q + 2]), acc);
				}
			}
		}
		if (t0 < 0)
		{
#pragma unroll
			for (int q = 0; q < R; q++)
			{
				const int row = row0 + q * S;
				y[row] = sum[q];
				epi.row(row, sum[q], __ldg(x + row), acc);
			}
		}
	}
	else if (masked && p >= 0)
	{	// the chains of pattern p = sup, every row with its own presence mask
		const int nch = s_info[p].info & 255;
		const PatChain* ch = s_ch + p * maxch;
		double sum[R];
#pragma unroll
		for (int q = 0; q < R; q++) sum[q] = 0.0;
		for (int c = 0; c < nch; c++)
		{
			double v0, v1, v2; int off, m;
			pat_chain_load(ch + c, v0, v1, v2, off, m);
			const int xb = row0 + off;
			unsigned int b[R + 2];
#pragma unroll
			for (int q = 0; q < R; q++) b[q] = (unsigned int)(mk[q] >> (kPatChainLen * c)) & 7u;
			b[R] = 0u; b[R + 1] = 0u;
			double xl[R + 2];
#pragma unroll
			for (int u = 0; u < R + 2; u++)
			{
				const unsigned int need = (b[u] & 1u) | (u >= 1 ? (b[u - 1] & 2u) : 0u) | (u >= 2 ? (b[u - 2] & 4u) : 0u);
				xl[u] = need ? __ldg(x + (xb + u * S)) : 0.0;
			}
#pragma unroll
			for (int q = 0; q < R; q++)
			{
				if (b[q] & 1u) sum[q] = fma(v0, xl[q], sum[q]);
				if (b[q] & 2u) sum[q] = fma(v1, xl[q + 1], sum[q]);
				if (b[q] & 4u) sum[q] = fma(v2, xl[q + 2], sum[q]);
			}
		}
#pragma unroll
		for (int q = 0; q < R; q++)
		{
			const int row = row0 + q * S;
			if (row < n_rows)
			{
				y[row] = sum[q];
				epi.row(row, sum[q], __ldg(x + row), acc);
			}
		}
	}
	else
	{
		for (int q = 0; q < R; q++)
		{
			const int row = row0 + q * S;
			if (row >= n_rows) break;
			const int pq = (int)pat[row];
			const int nch = s_info[pq].info & 255;
			const PatChain* ch = s_ch + pq * maxch;
			const double* xr = x + row;
			double sum = 0.0;
			for (int c = 0; c < nch; c++)
			{
				const int off = ch[c].off, m = ch[c].m & 3;
				for (int t = 0; t < m; t++) sum = fma(ch[c].v[t], __ldg(xr + off + t * S), sum);
			}
			y[row] = sum;
			epi.row(row, sum, __ldg(xr), acc);
		}
	}
}

// chains | info of all patterns -> shared memory (one device array, 16-byte words)
__device__ __forceinline__ void pat_tables_to_smem(const CsrDev<double>& A, unsigned char* smem_tables)
{
	const double2* g = reinterpret_cast<const double2*>(A.pat_chain);
	double2* s = reinterpret_cast<double2*>(smem_tables);
	const int n16 = 2 * A.n_pat * A.pat_maxch + A.n_pat;
	for (int i = threadIdx.x; i < n16; i += blockDim.x) s[i] = g[i];
}

template <class Epi>
__global__ void __launch_bounds__(kThreads, 3) k_spmv_pat(CsrDev<double> A, const double* __restrict__ x, double* __restrict__ y, Epi epi_in,
	DevState* st, double* partials)
{
	pdl_enter();
	if (st_done(st)) return;
	extern __shared__ __align__(128) unsigned char smem[];
	const PatChain* s_ch = reinterpret_cast<const PatChain*>(smem);
	const PatInfo* s_info = reinterpret_cast<const PatInfo*>(smem + (size_t)A.n_pat * A.pat_maxch * sizeof(PatChain));
	pat_tables_to_smem(A, smem);
	__syncthreads();

	Epi epi = epi_in;
	epi.begin(st);
	double acc[Epi::NRED > 0 ? Epi::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Epi::NRED > 0 ? Epi::NRED : 1); r++) acc[r] = 0.0;

	const int S = A.pat_stride, nib = A.pat_nib, n_items = A.pat_items;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	constexpr int WPB = kThreads / 32;
	// the warps of a block take the items of a round in rotation: the items at the ends of a grid line (mixed patterns,
	// slower) do not always land on the same two warps.  A thread's pattern byte (the pattern its R rows share, 255 = look at
	// the rows) is loaded one round ahead.
	auto item_of = [&](int round) { return (blockIdx.x + round * (int)gridDim.x) * WPB + ((warp + round) & (WPB - 1)); };
	int it = item_of(0);
	int tcode = it < n_items ? (int)A.pat_thread[(size_t)it * 32 + lane] : 255;
	for (int round = 0; (blockIdx.x + round * (int)gridDim.x) * WPB < n_items; round++)
	{
		const int it_next = item_of(round + 1);
		const int tcode_next = it_next < n_items ? (int)A.pat_thread[(size_t)it_next * 32 + lane] : 255;
		if (it < n_items)
		{
			const int a = it / nib, ib = it - a * nib;
			const int i = ib * 32 + lane;
			pat_item_ldg<Epi>(A.pat, x, y, s_ch, s_info, A.pat_maxch, S, A.n_rows, a * kPatRows * S + i, tcode, i < S, epi, acc);
		}
		it = it_next; tcode = tcode_next;
	}
	if (Epi::NRED > 0)
	{
		double tot[Epi::NRED > 0 ? Epi::NRED : 1];
		if (grid_reduce<(Epi::NRED > 0 ? Epi::NRED : 1)>(acc, partials, &st->ticket, tot))
		{
			if (st->multi) { if (reduce_across_ranks(st, tot, Epi::NRED)) epi.finish(st, tot); }
			else if ((threadIdx.x & 31) == 0) epi.finish(st, tot);
		}
	}
}

// global -> shared bulk copy without a cache hint (x is re-read by the neighbouring blocks: normal L2 residency)
__device__ __forceinline__ void bulk_g2s_plain(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// the marching kernel (paragraph 4).  WX = warps along the row index (8: S % 256 == 0, 4: S % 128 == 0), WY = 8 / WX warps
// along S.  Block item (Ab, ibb): warp (wq, wi) owns the rows ((Ab WY + wq) R + q) S + 32 (ibb WX + wi) + lane.
// pat_bitem[16 bi + w] = the pattern all rows of warp w's item share (255 = mixed, 254 = the item has no rows),
// pat_bitem[16 bi + 8] = 1 when every existing row of the block item is a subset of the geometry pattern.
template <class Epi, int WX>
__global__ void __launch_bounds__(kSpmvThreads, 2) k_spmv_pat_march(CsrDev<double> A, PatMarch M, const double* __restrict__ x, double* __restrict__ y, Epi epi_in,
	DevState* st, double* partials)
{
	pdl_enter();
	if (st_done(st)) return;
	constexpr int R = kPatRows, WY = 8 / WX;
	constexpr int WD = 32 * WX + kPatSpan;            // values per window line
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) unsigned long long s_bar[2 * kPatMaxStages];   // full[stage] | empty[stage]
	// layout: window stages | chains | info
	const int NST = M.nst;
	const int stage_bytes = M.nlines * WD * 8;
	// ... | the geometry pattern's chains | info of all patterns (the chains of the other patterns stay in global memory:
	// only block items that leave the march — the ghost-coupled planes of a row block — read them)
	unsigned char* s_tab = smem + (size_t)NST * stage_bytes;
	const PatChain* g_ch = reinterpret_cast<const PatChain*>(A.pat_chain);
	const PatChain* s_ch = reinterpret_cast<const PatChain*>(s_tab);
	const PatInfo* s_info = reinterpret_cast<const PatInfo*>(s_tab + (size_t)A.pat_maxch * sizeof(PatChain));
	const int tid = threadIdx.x;
	{
		const double2* gc = reinterpret_cast<const double2*>(g_ch + (size_t)M.gpat * A.pat_maxch);
		const double2* gi = reinterpret_cast<const double2*>(g_ch + (size_t)A.n_pat * A.pat_maxch);
		double2* sc = reinterpret_cast<double2*>(s_tab);
		for (int i = tid; i < 2 * A.pat_maxch; i += blockDim.x) sc[i] = gc[i];
		for (int i = tid; i < A.n_pat; i += blockDim.x) sc[2 * A.pat_maxch + i] = gi[i];
	}
	if (tid == 0)
	{
		for (int s = 0; s < NST; s++) { mbar_init(smem_u32(&s_bar[s]), 1); mbar_init(smem_u32(&s_bar[kPatMaxStages + s]), kThreads / 32); }
		mbar_fence_init();
	}
	__syncthreads();

	Epi epi = epi_in;
	double acc[Epi::NRED > 0 ? Epi::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Epi::NRED > 0 ? Epi::NRED : 1); r++) acc[r] = 0.0;
	const int S = A.pat_stride, n_rows = A.n_rows, n_cols = A.n_cols;
	const int nibb = S / (32 * WX);
	const int lane = tid & 31;
	const int4* segs = reinterpret_cast<const int4*>(A.pat_segs);
	// windows loaded / consumed so far by this block, over all its segments: number jg = use * NST + slot lives in stage `slot`
	int slot = 0, use = 0;

	if (tid >= kThreads)
	{	// ---- producer warp: one window per item (+ G - 1 at the head of a segment) ----
		for (int sg = blockIdx.x; sg < M.n_segs; sg += gridDim.x)
		{
			const int4 seg = __ldg(segs + sg);
			const int n_win = seg.z + M.G - 1;
			// x index behind line 0, position 0 of window 0 of the segment (may be negative: clipped below)
			long long wbase = (long long)seg.x * WY * R * S + (long long)seg.y * (32 * WX) + M.o0;
			for (int j = 0; j < n_win; j++, wbase += M.S2)
			{
				if (use > 0) mbar_wait(smem_u32(&s_bar[kPatMaxStages + slot]), (uint32_t)((use - 1) & 1));
				const uint32_t full = smem_u32(&s_bar[slot]);
				const uint32_t dst0 = smem_u32(smem) + (uint32_t)(slot * stage_bytes);
				// clip every line to [0, n_cols): both ends even, so copies stay 16-byte aligned and sized
				uint32_t my_bytes = 0;
				for (int l = lane; l < M.nlines; l += 32)
				{
					const long long b0 = wbase + (long long)l * S;
					const long long lo = b0 < 0 ? 0 : b0, hi = b0 + WD > n_cols ? n_cols : b0 + WD;
					if (hi > lo) my_bytes += (uint32_t)(hi - lo) * 8u;
				}
				const uint32_t total = (uint32_t)__reduce_add_sync(0xffffffffu, my_bytes);
				if (lane == 0) mbar_expect_tx(full, total);
				__syncwarp();
				for (int l = lane; l < M.nlines; l += 32)
				{
					const long long b0 = wbase + (long long)l * S;
					const long long lo = b0 < 0 ? 0 : b0, hi = b0 + WD > n_cols ? n_cols : b0 + WD;
					if (hi > lo) bulk_g2s_plain(dst0 + (uint32_t)((l * WD + (int)(lo - b0)) * 8), x + lo, (uint32_t)(hi - lo) * 8u, full);
				}
				if (++slot == NST) { slot = 0; use++; }
			}
		}
	}
	else
	{	// ---- consumer warps ----
		epi.begin(st);
		const int warp = tid >> 5;
		const int wq = warp / WX, wi = warp % WX;
		const unsigned char* bitem = A.pat_bitem;
		const int gpat = M.gpat;
		const PatInfo ginfo = s_info[gpat];
		const int t0 = (ginfo.info >> 8) - 1, c_last = (ginfo.info & 255) - 1;
		const PatChain* ch = s_ch;
		const double* win0 = reinterpret_cast<const double*>(smem) + (wq * R) * WD + wi * 32 + lane;   // my line 0, my position, stage 0
		const int stage_dbl = stage_bytes / 8;
		for (int sg = blockIdx.x; sg < M.n_segs; sg += gridDim.x)
		{
			const int4 seg = __ldg(segs + sg);
			long long bi = (long long)seg.x * nibb + seg.y;
			const long long bi_step = (long long)M.dAb * nibb;
			int row0 = (seg.x * WY + wq) * R * S + (seg.y * WX + wi) * 32 + lane;
			// my thread's byte of k_spmv_pat's item (A, ib) = (Ab WY + wq, ibb WX + wi): index (A nib + ib) 32 + lane, nib = S / 32
			long long ti = ((long long)(seg.x * WY + wq) * (S / 32) + (seg.y * WX + wi)) * 32 + lane;
			const long long ti_step = (long long)M.dAb * WY * S;
			unsigned int code = (unsigned int)bitem[16 * bi + warp] | ((unsigned int)bitem[16 * bi + 8] << 8);
			if ((code & 255u) == 255u) code |= (unsigned int)A.pat_thread[ti] << 16;
			// the first item needs windows 0 .. G - 1: wait for the first G - 1 here, the last one inside the loop
			for (int j = 0; j < M.G - 1; j++)
			{
				const int sj = slot + j >= NST ? slot + j - NST : slot + j;
				mbar_wait(smem_u32(&s_bar[sj]), (uint32_t)((slot + j >= NST ? use + 1 : use) & 1));
			}
			for (int k = 0; k < seg.z; k++, bi += bi_step, row0 += M.S2, ti += ti_step)
			{
				unsigned int code_next = 0xfe;
				if (k + 1 < seg.z)
				{
					code_next = (unsigned int)bitem[16 * (bi + bi_step) + warp] | ((unsigned int)bitem[16 * (bi + bi_step) + 8] << 8);
					if ((code_next & 255u) == 255u) code_next |= (unsigned int)A.pat_thread[ti + ti_step] << 16;
				}
				{	// window k + G - 1, the last one this item needs
					const int j = M.G - 1, sj = slot + j >= NST ? slot + j - NST : slot + j;
					mbar_wait(smem_u32(&s_bar[sj]), (uint32_t)((slot + j >= NST ? use + 1 : use) & 1));
				}
				const int uni = code & 255;
				const int tcode = uni != 255 ? uni : (int)(code >> 16);   // the pattern my R rows share, 255 = mixed
				if (uni != 254)
				{
					if ((code >> 8) & 255u)
					{	// every row is a subset of the geometry pattern: its chains, read out of the plane windows
						unsigned long long pqs = 0ull, mk0;
						bool same = true;
						if (tcode != 255) mk0 = s_info[tcode].mask;
						else
						{	// my rows' pattern ids (one byte each, 255 = row beyond the matrix) + the first row's mask
							mk0 = 0ull;
#pragma unroll
							for (int q = 0; q < R; q++)
							{
								const int row = row0 + q * S;
								const int pq = row < n_rows ? (int)A.pat[row] : 255;
								const unsigned long long mq = pq != 255 ? s_info[pq].mask : 0ull;
								if (q == 0) mk0 = mq;
								same = same && mq == mk0;
								pqs |= (unsigned long long)pq << (8 * q);
							}
						}
						double sum[R];
#pragma unroll
						for (int q = 0; q < R; q++) sum[q] = 0.0;
						for (int g = 0; g < M.G; g++)
						{
							const int sp = slot + M.group_plane(g);
							const double* win = win0 + (sp >= NST ? sp - NST : sp) * stage_dbl;
							const int c_end = M.group_begin(g + 1);
							for (int c = M.group_begin(g); c < c_end; c++)
							{
								double v0, v1, v2; int off, m;
								pat_chain_load(ch + c, v0, v1, v2, off, m);
								const double* src = win + ((m >> 4) & 15) * WD + ((m >> 8) & 255);
								double xl[R + 2];
#pragma unroll
								for (int u = 0; u < R; u++) xl[u] = src[u * WD];
								xl[R] = (m & 3) > 1 ? src[R * WD] : 0.0;           // the window has WY R + max(shift + m - 1) lines
								xl[R + 1] = (m & 3) > 2 ? src[(R + 1) * WD] : 0.0;
								if (same)
								{
									const unsigned int pres = (unsigned int)(mk0 >> (kPatChainLen * c)) & 7u;
									if (pres & 1u)
									{
#pragma unroll
										for (int q = 0; q < R; q++) sum[q] = fma(v0, xl[q], sum[q]);
									}
									if (pres & 2u)
									{
#pragma unroll
										for (int q = 0; q < R; q++) sum[q] = fma(v1, xl[q + 1], sum[q]);
									}
									if (pres & 4u)
									{
#pragma unroll
										for (int q = 0; q < R; q++) sum[q] = fma(v2, xl[q + 2], sum[q]);
									}
								}
								else
								{
#pragma unroll
									for (int q = 0; q < R; q++)
									{
										const int pq = (int)(pqs >> (8 * q)) & 255;
										const unsigned int pres = pq != 255 ? ((unsigned int)(s_info[pq].mask >> (kPatChainLen * c)) & 7u) : 0u;
										if (pres & 1u) sum[q] = fma(v0, xl[q], sum[q]);
										if (pres & 2u) sum[q] = fma(v1, xl[q + 1], sum[q]);
										if (pres & 4u) sum[q] = fma(v2, xl[q + 2], sum[q]);
									}
								}
								if (c == c_last && t0 >= 0)
								{	// the diagonal's chain is the LAST chain of the last group: the sums are complete and x[row] is in xl
#pragma unroll
									for (int q = 0; q < R; q++)
									{
										const int row = row0 + q * S;
										if (tcode != 255 || row < n_rows)
										{
											y[row] = sum[q];
											epi.row(row, sum[q], t0 == 0 ? xl[q] : (t0 == 1 ? xl[q + 1] : xl[q + 2]), acc);
										}
									}
								}
							}
						}
						if (t0 < 0)
						{
#pragma unroll
							for (int q = 0; q < R; q++)
							{
								const int row = row0 + q * S;
								if (uni != 255 || row < n_rows)
								{
									y[row] = sum[q];
									epi.row(row, sum[q], __ldg(x + row), acc);
								}
							}
						}
					}
					else pat_item_ldg<Epi>(A.pat, x, y, g_ch, s_info, A.pat_maxch, S, n_rows, row0, tcode, true, epi, acc);   // e.g. the ghost-coupled planes of a row block
				}
				__syncwarp();
				if (lane == 0) mbar_arrive(smem_u32(&s_bar[kPatMaxStages + slot]));   // window k is not needed by item k + 1
				if (++slot == NST) { slot = 0; use++; }
				code = code_next;
			}
			// the G - 1 windows behind the last item of the segment
			for (int j = 0; j < M.G - 1; j++)
			{
				if (lane == 0) mbar_arrive(smem_u32(&s_bar[kPatMaxStages + slot]));
				if (++slot == NST) { slot = 0; use++; }
			}
		}
	}
	if (Epi::NRED > 0)
	{
		double tot[Epi::NRED > 0 ? Epi::NRED : 1];
		if (grid_reduce<(Epi::NRED > 0 ? Epi::NRED : 1)>(acc, partials, &st->ticket, tot))
		{
			if (st->multi) { if (reduce_across_ranks(st, tot, Epi::NRED)) epi.finish(st, tot); }
			else if ((threadIdx.x & 31) == 0) epi.finish(st, tot);
		}
	}
}

// ---- the box kernel (paragraph 5 of the overview above) ------------------------------------------------------------------
// Geometry pattern = a dense box: G planes x (dx = -1, 0, +1) x (dy = -1, 0, +1) around even centres (pat_host.h:
// pat_plan_box).  A thread owns an aligned PAIR of rows on each of R lines: per plane and window line it loads the pair
// (one 16-byte load) and the value on either side, and 18 fma consume them — 3.75 loads and 3 load instructions per 18
// products, against 10 loads per 24 products in k_spmv_pat.  The 27 coefficients are kernel parameters: the fma read them
// straight out of the constant bank, no register, no table read.  Everything is unrolled: no index arithmetic per chain.
// Boundary rows are the box minus whole slices, so a byte of flags per thread says what to leave out: the value left /
// right of the pair (dx), the first / last window line (dy: it serves only the thread's first / last row), a whole plane
// (dz).  255 = rows that are no such sub-box (or lie beyond the matrix): both columns go through pat_item_ldg; try_patterns
// takes this kernel only when at most a tenth of the warps hold such a thread.
// Measured at 27-point 256^3 (profiles/README_r02.md): 0.083 ms = 0.52 of the measured HBM peak on the 17 bytes per row,
// 916 instructions per 512 rows (448 of them fma), LSU data pipe 75 % (the two side values are 8-byte loads with a 16-byte lane
// stride: 4 wavefronts each, like the pair), 124 registers, 2 blocks per SM.  Handing the side values over by warp shuffles
// instead costs more than it saves: 0.208 ms with a shuffle behind every load, 0.176 ms with all pair loads of a plane issued first
// (tried twice, dropped).
struct PatBox {
	int G = 0;
	int center[3] = {0, 0, 0};
	double coef[3][3][3] = {};   // [plane][dx + 1][line j]
};
constexpr int kBoxDropL = 1, kBoxDropR = 2, kBoxDropLow = 4, kBoxDropHigh = 8, kBoxDropG0 = 16;

template <class Epi>
__global__ void __launch_bounds__(kThreads, 2) k_spmv_pat_box(CsrDev<double> A, PatBox B, const double* __restrict__ x, double* __restrict__ y, Epi epi_in,
	DevState* st, double* partials)
{
	pdl_enter();
	if (st_done(st)) return;
	constexpr int R = kPatRows;
	extern __shared__ __align__(128) unsigned char smem[];
	const PatChain* s_ch = reinterpret_cast<const PatChain*>(smem);   // the chain tables serve the rows that are no sub-box
	const PatInfo* s_info = reinterpret_cast<const PatInfo*>(smem + (size_t)A.n_pat * A.pat_maxch * sizeof(PatChain));
	pat_tables_to_smem(A, smem);
	__syncthreads();

	Epi epi = epi_in;
	epi.begin(st);
	double acc[Epi::NRED > 0 ? Epi::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Epi::NRED > 0 ? Epi::NRED : 1); r++) acc[r] = 0.0;

	const int S = A.pat_stride, n_rows = A.n_rows;
	const int nib = (S + 63) / 64;                       // warp items per R lines: 64 rows of every line
	const int n_items = A.pat_items / A.pat_nib * nib;   // pat_items = groups of R lines x pat_nib
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	constexpr int WPB = kThreads / 32;
	auto item_of = [&](int round) { return (blockIdx.x + round * (int)gridDim.x) * WPB + ((warp + round) & (WPB - 1)); };
	int it = item_of(0);
	int flags = it < n_items ? (int)A.pat_box_flags[(size_t)it * 32 + lane] : 255;
	for (int round = 0; (blockIdx.x + round * (int)gridDim.x) * WPB < n_items; round++)
	{
		const int it_next = item_of(round + 1);
		const int flags_next = it_next < n_items ? (int)A.pat_box_flags[(size_t)it_next * 32 + lane] : 255;
		if (it < n_items)
		{
			const int a = it / nib, ib = it - a * nib;
			const int i = ib * 64 + 2 * lane;
			const int row0 = a * R * S + i;   // even: S is even
			if (i < S)
			{
				if (flags != 255)
				{
					double s0[R], s1[R];
#pragma unroll
					for (int q = 0; q < R; q++) { s0[q] = 0.0; s1[q] = 0.0; }
#pragma unroll
					for (int g = 0; g < 3; g++)
					{
						if (g < B.G && !(flags & (kBoxDropG0 << g)))
						{
							const int xg = row0 + B.center[g];
#pragma unroll
							for (int u = 0; u < R + 2; u++)
							{
								double xm = 0.0, x0 = 0.0, x1 = 0.0, xp = 0.0;
								const bool line_on = !((u == 0 && (flags & kBoxDropLow)) || (u == R + 1 && (flags & kBoxDropHigh)));
								if (line_on)
								{
									const double* pl = x + (xg + u * S);
									const double2 c = __ldg(reinterpret_cast<const double2*>(pl));
									x0 = c.x; x1 = c.y;
									if (!(flags & kBoxDropL)) xm = __ldg(pl - 1);
									if (!(flags & kBoxDropR)) xp = __ldg(pl + 2);
								}
#pragma unroll
								for (int j = 0; j < 3; j++)
								{
									const int q = u - j;
									if (q >= 0 && q < R)
									{
										s0[q] = fma(B.coef[g][0][j], xm, s0[q]); s0[q] = fma(B.coef[g][1][j], x0, s0[q]); s0[q] = fma(B.coef[g][2][j], x1, s0[q]);
										s1[q] = fma(B.coef[g][0][j], x0, s1[q]); s1[q] = fma(B.coef[g][1][j], x1, s1[q]); s1[q] = fma(B.coef[g][2][j], xp, s1[q]);
									}
								}
							}
						}
					}
#pragma unroll
					for (int q = 0; q < R; q++)
					{
						const int row = row0 + q * S;
						const double2 xc = __ldg(reinterpret_cast<const double2*>(x + row));   // dead code for epilogues that ignore x[row]
						*reinterpret_cast<double2*>(y + row) = make_double2(s0[q], s1[q]);
						epi.row(row, s0[q], xc.x, acc);
						epi.row(row + 1, s1[q], xc.y, acc);
					}
				}
				else
				{	// no sub-box: the two columns one after the other through the chain tables (k_spmv_pat's thread item (a, i / 32, i % 32))
					for (int e = 0; e < 2; e++)
					{
						const int tcode = (int)A.pat_thread[(size_t)a * A.pat_nib * 32 + i + e];
						pat_item_ldg<Epi>(A.pat, x, y, s_ch, s_info, A.pat_maxch, S, n_rows, row0 + e, tcode, true, epi, acc);
					}
				}
			}
		}
		it = it_next; flags = flags_next;
	}
	if (Epi::NRED > 0)
	{
		double tot[Epi::NRED > 0 ? Epi::NRED : 1];
		if (grid_reduce<(Epi::NRED > 0 ? Epi::NRED : 1)>(acc, partials, &st->ticket, tot))
		{
			if (st->multi) { if (reduce_across_ranks(st, tot, Epi::NRED)) epi.finish(st, tot); }
			else if ((threadIdx.x & 31) == 0) epi.finish(st, tot);
		}
	}
}

inline size_t pat_table_bytes(int n_pat, int maxch) { return (size_t)n_pat * maxch * sizeof(PatChain) + (size_t)n_pat * sizeof(PatInfo); }
// the marching kernel: window stages + the geometry pattern's chains + info of all patterns.  Two blocks fit an SM (227 KB,
// 1 KB reserved and about 1.3 KB of static shared memory per block) up to kPatMarchSmemFor2 each
inline size_t pat_march_smem_bytes(int nst, int nlines, int wx, int n_pat, int maxch)
{
	return (size_t)nst * nlines * (32 * wx + kPatSpan) * 8 + (size_t)maxch * sizeof(PatChain) + (size_t)n_pat * sizeof(PatInfo);
}
constexpr int kPatMarchSmemFor2 = 111 * 1024;

template <class Epi, int WX>
inline void launch_spmv_pat_march(const CsrDev<double>& A, const double* x, double* y, const Epi& epi, DevState* st, double* partials, cudaStream_t s)
{
	const PatMarch& M = *A.pat_march;
	const size_t smem = pat_march_smem_bytes(M.nst, M.nlines, WX, A.n_pat, A.pat_maxch);
	auto kern = k_spmv_pat_march<Epi, WX>;
	static PerDeviceOnce once;
	if (once.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	const int limit = spmv_grid_limit(smem <= (size_t)kPatMarchSmemFor2 ? 2 : 1);
	int grid = M.n_segs < limit ? M.n_segs : limit;
	if (grid < 1) grid = 1;
	launch_k(kern, grid, kSpmvThreads, smem, s, A, M, x, y, epi, st, partials);
}

template <class Epi>
inline void launch_spmv_pat(const CsrDev<double>& A, const double* x, double* y, const Epi& epi, DevState* st, double* partials, cudaStream_t s)
{
	if (A.pat_box && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0)
	{	// 16-byte loads of x and stores of y: the plan guarantees even row indices
		const size_t smem = pat_table_bytes(A.n_pat, A.pat_maxch);
		auto kern = k_spmv_pat_box<Epi>;
		static PerDeviceOnce once;
		if (once.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kPatMaxChains * sizeof(PatChain) + 256 * sizeof(PatInfo)));
		const int n_items = A.pat_items / A.pat_nib * ((A.pat_stride + 63) / 64);
		const int n_blocks = (n_items + kThreads / 32 - 1) / (kThreads / 32);
		const int limit = spmv_grid_limit(2);
		int grid = n_blocks < limit ? n_blocks : limit;
		if (grid < 1) grid = 1;
		launch_k(kern, grid, kThreads, smem, s, A, *A.pat_box, x, y, epi, st, partials);
		return;
	}
	if (A.pat_march && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
	{	// TMA bulk copies need 16-byte aligned sources; the plan already guarantees even row bases and an even vector length
		if (A.pat_march->wx == 8) launch_spmv_pat_march<Epi, 8>(A, x, y, epi, st, partials, s);
		else launch_spmv_pat_march<Epi, 4>(A, x, y, epi, st, partials, s);
		return;
	}
	const size_t smem = pat_table_bytes(A.n_pat, A.pat_maxch);
	auto kern = k_spmv_pat<Epi>;
	static PerDeviceOnce once;
	if (once.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kPatMaxChains * sizeof(PatChain) + 256 * sizeof(PatInfo)));
	const int n_blocks = (A.pat_items + kThreads / 32 - 1) / (kThreads / 32);
	const int limit = spmv_grid_limit(3);
	int grid = n_blocks < limit ? n_blocks : limit;
	if (grid < 1) grid = 1;
	launch_k(kern, grid, kThreads, smem, s, A, x, y, epi, st, partials);
}

// Adaptor: run an SpMV epilogue as a plain vector kernel over an already computed y (user-callback operators).
template <class T, class Epi>
struct RowEpilogueOp {
	static constexpr int NRED = Epi::NRED;
	static constexpr int W = 1;
	Epi epi; const T* x; const T* y;
	__device__ bool active(const DevState*) const { return true; }
	__device__ void begin(const DevState* st) { epi.begin(st); }
	template <int V> __device__ void elem(size_t i, double* acc) { epi.row((int)i, y[i], x[i], acc); }
	__device__ void finish(DevState* st, const double* tot) { epi.finish(st, tot); }
};

// epilogue that does nothing (plain y = A x)
template <class T>
struct EpiNone {
	static constexpr int NRED = 0;
	static constexpr bool ACTIVE = false;   // nothing to run after a user-callback SpMV
	__device__ void begin(const DevState*) {}
	__device__ void row(int, T, T, double*) {}
	__device__ void finish(DevState*, const double*) {}
};


template <class T, int LPR, bool CONJ, class Epi>
inline void launch_spmv_lpr(const CsrDev<T>& A, const T* x, T* y, const Epi& epi, DevState* st, double* partials, cudaStream_t s)
{
	const int n_chunks = (A.n_tiles + A.chunk - 1) / A.chunk;
	const int limit = spmv_grid_limit(TileCfg<T>::CTAS);
	int grid = n_chunks < limit ? n_chunks : limit;
	if (grid < 1) grid = 1;
	if (A.n_interior >= 0)
	{
		static PerDeviceOnce once;   // per instantiation
		auto kern = k_spmv<T, LPR, CONJ, Epi, true>;
		if (once.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, StageCfg<T>::TOTAL);
		launch_k(kern, grid, kSpmvThreads, StageCfg<T>::TOTAL, s, A, x, y, epi, st, partials);
	}
	else
	{
		static PerDeviceOnce once;
		auto kern = k_spmv<T, LPR, CONJ, Epi, false>;
		if (once.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, StageCfg<T>::TOTAL);
		launch_k(kern, grid, kSpmvThreads, StageCfg<T>::TOTAL, s, A, x, y, epi, st, partials);
	}
}

template <int LPR, class Epi>
inline void launch_spmv_dict_lpr(const CsrDev<double>& A, const double* x, double* y, const Epi& epi, DevState* st, double* partials, cudaStream_t s)
{
	static PerDeviceOnce once;
	auto kern = k_spmv_dict<LPR, Epi>;
	if (once.first()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DictStage::TOTAL);
	const int n_chunks = (A.n_dtiles + A.dchunk - 1) / A.dchunk;
	const int limit = spmv_grid_limit(kSpmvCtasPerSm);
	int grid = n_chunks < limit ? n_chunks : limit;
	if (grid < 1) grid = 1;
	launch_k(kern, grid, kSpmvThreads, DictStage::TOTAL, s, A, x, y, epi, st, partials);
}

template <class Epi>
inline void launch_spmv_dict(const CsrDev<double>& A, const double* x, double* y, const Epi& epi, DevState* st, double* partials, cudaStream_t s)
{
	switch (A.dlpr)
	{
		case 1: launch_spmv_dict_lpr<1, Epi>(A, x, y, epi, st, partials, s); break;
		case 2: launch_spmv_dict_lpr<2, Epi>(A, x, y, epi, st, partials, s); break;
		case 4: launch_spmv_dict_lpr<4, Epi>(A, x, y, epi, st, partials, s); break;
		case 8: launch_spmv_dict_lpr<8, Epi>(A, x, y, epi, st, partials, s); break;
		case 16: launch_spmv_dict_lpr<16, Epi>(A, x, y, epi, st, partials, s); break;
		default: launch_spmv_dict_lpr<32, Epi>(A, x, y, epi, st, partials, s); break;
	}
}

// launch with the lanes-per-row variant recorded in the handle
template <class T, bool CONJ, class Epi>
inline void launch_spmv(const CsrDev<T>& A, const T* x, T* y, const Epi& epi, DevState* st, double* partials, cudaStream_t s)
{
	if constexpr (std::is_same<T, double>::value && !CONJ)
	{
		if (A.pat) { launch_spmv_pat<Epi>(A, x, y, epi, st, partials, s); return; }      // row-pattern copy: 1 B per ROW
		if (A.code) { launch_spmv_dict<Epi>(A, x, y, epi, st, partials, s); return; }   // dictionary copy: 2 B per entry
	}
	switch (A.lpr)
	{
		case 1: launch_spmv_lpr<T, 1, CONJ, Epi>(A, x, y, epi, st, partials, s); break;
		case 2: launch_spmv_lpr<T, 2, CONJ, Epi>(A, x, y, epi, st, partials, s); break;
		case 4: launch_spmv_lpr<T, 4, CONJ, Epi>(A, x, y, epi, st, partials, s); break;
		case 8: launch_spmv_lpr<T, 8, CONJ, Epi>(A, x, y, epi, st, partials, s); break;
		case 16: launch_spmv_lpr<T, 16, CONJ, Epi>(A, x, y, epi, st, partials, s); break;
		default: launch_spmv_lpr<T, 32, CONJ, Epi>(A, x, y, epi, st, partials, s); break;
	}
}

}  // namespace lcgb200
