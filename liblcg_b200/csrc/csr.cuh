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
	const unsigned char* pat = nullptr; const unsigned char* pat_item = nullptr; const void* pat_info = nullptr; const void* pat_chain = nullptr;
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
// the number of gathers the load/store pipe has to serve (round 1: 27 per row, LSU data pipe 84 % busy, 0.2 of HBM).
//
// Round 2: gathers shared in REGISTERS.  The entries of a pattern are grouped into CHAINS: up to kPatChainLen offsets in
// arithmetic progression with the matrix-wide stride S (for a 3-D stencil S = nx: the entries (dx, dz) fixed, dy = -1, 0,
// +1).  A thread owns R rows that are S apart (row, row + S, ..., row + (R-1) S: a column of R grid points in y), so the
// x values one chain needs for all R rows are R + m - 1 loads instead of R * m — x[row + off + u S] serves row q = u - t
// through chain entry t.  27-point stencil, R = 8: 9 chains x 10 loads for 8 rows = 11.25 gathers per row instead of 27,
// and one 32-byte table read per chain instead of one per entry.  Lanes of a warp sit on 32 consecutive rows, so every
// gather is 32 consecutive doubles (2-3 lines), every store a full 256-byte line.
//
// Work item of a warp = R x 32 rows: rows (A R + q) S + 32 ib + lane.  A byte per ITEM (pat_item) says whether all its
// rows exist and share one pattern (the rule inside a stencil): then the warp does not even read the per-row ids and all
// table reads are broadcasts.  Otherwise each thread looks at its own R rows: same pattern -> the chain path with a
// per-lane pattern; patterns that are SUBSETS of one longer pattern (a column of grid points that starts on a face: the
// face row is the interior row minus the entries that leave the grid; pat_host.h: pat_build_masks) -> the longer pattern's
// chains with a 64-bit presence mask per row, loads still shared; anything else row by row through the same chain table.
// Entries are accumulated chain by chain with fma (a different order from the CSR row order: y agrees with the plain
// copy to rounding, not bitwise).
constexpr int kPatRows = 8;            // R: rows (S apart) a thread computes together
constexpr int kPatChainLen = 3;        // entries per chain (v[3])
constexpr int kPatMaxChains = 3072;    // chain table entries held in shared memory (96 KB)
constexpr int kPatDefaultStride = 256; // S when no pair of offsets repeats (all chains have one entry)
struct __align__(16) PatChain { double v[kPatChainLen]; int off; int m; };   // 32 bytes: two 128-bit shared-memory reads
static_assert(sizeof(PatChain) == 32, "PatChain is read as two 16-byte words");
// per pattern: presence mask over the chains of pattern `sup` (bit 3 c + t), chains | (index of offset 0 in chain 0, +1) << 8
struct __align__(16) PatInfo { unsigned long long mask; int info; int sup; };

__device__ __forceinline__ void pat_chain_load(const PatChain* c, double& v0, double& v1, double& v2, int& off, int& m)
{
	const double2 a = reinterpret_cast<const double2*>(c)[0];
	const double2 b = reinterpret_cast<const double2*>(c)[1];
	v0 = a.x; v1 = a.y; v2 = b.x;
	const long long om = __double_as_longlong(b.y);
	off = (int)(om & 0xffffffffll); m = (int)(om >> 32);
}

template <class Epi>
__global__ void __launch_bounds__(kThreads, 3) k_spmv_pat(CsrDev<double> A, const double* __restrict__ x, double* __restrict__ y, Epi epi_in,
	DevState* st, double* partials)
{
	pdl_enter();
	if (st_done(st)) return;
	constexpr int R = kPatRows;
	extern __shared__ __align__(128) unsigned char smem[];
	PatChain* s_ch = reinterpret_cast<PatChain*>(smem);
	const int n_ch = A.n_pat * A.pat_maxch;
	PatInfo* s_info = reinterpret_cast<PatInfo*>(smem + (size_t)n_ch * sizeof(PatChain));
	{
		const double2* g = reinterpret_cast<const double2*>(A.pat_chain);
		double2* s = reinterpret_cast<double2*>(smem);
		for (int i = threadIdx.x; i < 2 * n_ch; i += blockDim.x) s[i] = g[i];
		const double2* gi = reinterpret_cast<const double2*>(A.pat_info);
		double2* si = reinterpret_cast<double2*>(s_info);
		for (int i = threadIdx.x; i < A.n_pat; i += blockDim.x) si[i] = gi[i];
	}
	__syncthreads();

	Epi epi = epi_in;
	epi.begin(st);
	double acc[Epi::NRED > 0 ? Epi::NRED : 1];
#pragma unroll
	for (int r = 0; r < (Epi::NRED > 0 ? Epi::NRED : 1); r++) acc[r] = 0.0;

	const int S = A.pat_stride, nib = A.pat_nib, n_items = A.pat_items, maxch = A.pat_maxch;
	const long long n_rows = A.n_rows;
	const size_t Ss = (size_t)S;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	constexpr int WPB = kThreads / 32;
	for (int it = blockIdx.x * WPB + warp; it < n_items; it += gridDim.x * WPB)
	{
		const int a = it / nib, ib = it - a * nib;
		const int i = ib * 32 + lane;
		if (i >= S) continue;
		const long long row0 = (long long)a * R * S + i;
		const int uni = (int)A.pat_item[it];   // warp-uniform: pattern id shared by the whole item, 255 = mixed
		int p = uni;
		bool chained = true, masked = false;
		unsigned long long mk[R];
		if (uni == 255)
		{
			int p0 = -1, sup = -1;
			masked = true;
#pragma unroll
			for (int q = 0; q < R; q++)
			{
				const long long row = row0 + (long long)q * S;
				const int pq = row < n_rows ? (int)A.pat[row] : -1;
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
			const double* xr = x + row0;
			double sum[R], xc[R];
#pragma unroll
			for (int q = 0; q < R; q++) { sum[q] = 0.0; xc[q] = 0.0; }
			for (int c = 0; c < nch; c++)
			{
				double v0, v1, v2; int off, m;
				pat_chain_load(ch + c, v0, v1, v2, off, m);
				const double* xb = xr + off;
				double xl[R + 2];
#pragma unroll
				for (int u = 0; u < R; u++) xl[u] = __ldg(xb + u * Ss);
				xl[R] = m > 1 ? __ldg(xb + R * Ss) : 0.0;
				xl[R + 1] = m > 2 ? __ldg(xb + (R + 1) * Ss) : 0.0;
				if (c == 0)
				{	// chain 0 holds the diagonal when the row has one: keep x[row] for the epilogue
#pragma unroll
					for (int q = 0; q < R; q++) xc[q] = t0 == 0 ? xl[q] : (t0 == 1 ? xl[q + 1] : xl[q + 2]);
				}
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
			}
			if (t0 < 0)
			{
#pragma unroll
				for (int q = 0; q < R; q++) xc[q] = __ldg(xr + q * Ss);
			}
#pragma unroll
			for (int q = 0; q < R; q++)
			{
				const long long row = row0 + (long long)q * S;
				y[row] = sum[q];
				epi.row((int)row, sum[q], xc[q], acc);
			}
		}
		else if (masked && p >= 0)
		{	// the chains of pattern p = sup, every row with its own presence mask
			const int nch = s_info[p].info & 255;
			const PatChain* ch = s_ch + p * maxch;
			const double* xr = x + row0;
			double sum[R];
#pragma unroll
			for (int q = 0; q < R; q++) sum[q] = 0.0;
			for (int c = 0; c < nch; c++)
			{
				double v0, v1, v2; int off, m;
				pat_chain_load(ch + c, v0, v1, v2, off, m);
				const double* xb = xr + off;
				unsigned int b[R + 2];
#pragma unroll
				for (int q = 0; q < R; q++) b[q] = (unsigned int)(mk[q] >> (kPatChainLen * c)) & 7u;
				b[R] = 0u; b[R + 1] = 0u;
				double xl[R + 2];
#pragma unroll
				for (int u = 0; u < R + 2; u++)
				{
					const unsigned int need = (b[u] & 1u) | (u >= 1 ? (b[u - 1] & 2u) : 0u) | (u >= 2 ? (b[u - 2] & 4u) : 0u);
					xl[u] = need ? __ldg(xb + u * Ss) : 0.0;
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
				const long long row = row0 + (long long)q * S;
				if (row < n_rows)
				{
					y[row] = sum[q];
					epi.row((int)row, sum[q], __ldg(x + row), acc);
				}
			}
		}
		else
		{
			for (int q = 0; q < R; q++)
			{
				const long long row = row0 + (long long)q * S;
				if (row >= n_rows) break;
				const int pq = (int)A.pat[row];
				const int nch = s_info[pq].info & 255;
				const PatChain* ch = s_ch + pq * maxch;
				const double* xr = x + row;
				double sum = 0.0;
				for (int c = 0; c < nch; c++)
				{
					const int off = ch[c].off, m = ch[c].m;
					for (int t = 0; t < m; t++) sum = fma(ch[c].v[t], __ldg(xr + off + t * Ss), sum);
				}
				y[row] = sum;
				epi.row((int)row, sum, __ldg(xr), acc);
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

template <class Epi>
inline void launch_spmv_pat(const CsrDev<double>& A, const double* x, double* y, const Epi& epi, DevState* st, double* partials, cudaStream_t s)
{
	const size_t smem = (size_t)A.n_pat * A.pat_maxch * sizeof(PatChain) + (size_t)A.n_pat * sizeof(PatInfo);
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
