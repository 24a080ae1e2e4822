// capi.cu — the extern "C" boundary (include/lcgb200.h): CSR handle management, parameter validation in the
// reference's order, host<->device staging of m/B as the reference does it (lcg_cuda.cu:110-111,210), and the
// adaptors that let user Ax/Mx callbacks (cuSPARSE descriptors) drive the same engine.
#include "solvers.cuh"
#include "ic0_host.h"
#include "pat_host.h"
#include <dlfcn.h>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <type_traits>

using namespace lcgb200;

namespace lcgb200 {
const char* last_error();
}

// ------------------------------------------------------------------------------------------------ helpers
namespace {

struct ApiFailure { int code; };

template <class T> T* dev_alloc(size_t count)
{
	T* p = nullptr;
	LCG_CUDA_CHECK(cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T)));
	return p;
}

// lanes per row: the smallest power of two that leaves each lane at most ~7 entries of a mean-length row
int pick_lpr(long long nnz, long long rows)
{
	const double avg = rows > 0 ? (double)nnz / (double)rows : 0.0;
	int lpr = 1;
	while (lpr < 32 && avg > 7.0 * lpr) lpr *= 2;
	return lpr;
}

// rows per tile: a multiple of the block's row groups (kThreads / lpr) so that the compute phase of a typical
// tile has no ragged last pass, as many multiples as the staging buffer holds at the mean row length
int pick_tile_rows(long long nnz, long long rows, int lpr, int tile_nnz)
{
	const int ng = kThreads / lpr;
	const double avg = rows > 0 ? (double)nnz / (double)rows : 1.0;
	int passes = (int)((double)tile_nnz / (std::max(avg, 1.0) * ng));
	if (passes < 1) passes = 1;
	int cap = ng * passes;
	if (cap > kTileRows) cap = (kTileRows / ng) * ng;
	if (cap < 1) cap = kTileRows;
	return cap;
}

// tiles per chunk: a CTA walks `chunk` consecutive tiles (~256 rows: one grid line of a 256^3 stencil; measured best on B200 with 4 CTAs per SM, profiles/sweep_r01.txt) before jumping ahead by the grid stride
int pick_chunk(int tile_rows)
{
	static const char* env = getenv("LCGB200_SPMV_CHUNK_ROWS");
	const int target = env ? std::max(1, atoi(env)) : 256;
	return std::max(1, target / std::max(tile_rows, 1));
}

// greedy row tiles: consecutive rows (at most rows_cap) whose non-zeros, counted from the `align`-aligned start, fit one stage
void build_tiles(const int* rp, int n_rows, int tile_nnz, int rows_cap, std::vector<int4>& tiles, int align = 4)
{
	tiles.clear();
	int r = 0;
	while (r < n_rows)
	{
		const int k0 = rp[r] & ~(align - 1);
		int r1 = r;
		while (r1 < n_rows && r1 - r < rows_cap && rp[r1 + 1] - k0 <= tile_nnz) r1++;
		if (r1 == r) r1 = r + 1;   // a single row longer than the buffer: the kernel streams it from global memory
		tiles.push_back(make_int4(r, r1, k0, rp[r1]));
		r = r1;
	}
}

__global__ void k_diag_real(int n, const int* rp, const int* ci, const double* v, double* d)
{	// first entry with col == row, like lcg_smDcsr_get_diagonal_device (algebra_cuda.cu:40-57); 0 when absent
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double out = 0.0;
	for (int k = rp[i]; k < rp[i + 1]; k++) if (ci[k] == i) { out = v[k]; break; }
	d[i] = out;
}
__global__ void k_diag_cplx(int n, const int* rp, const int* ci, const double2* v, double2* d)
{	// lcg_complex_cuda.cu:46-63
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double2 out = make_double2(0.0, 0.0);
	for (int k = rp[i]; k < rp[i + 1]; k++) if (ci[k] == i) { out = v[k]; break; }
	d[i] = out;
}

// *fail = 1 if any column index lies outside [0, n_cols): a bad index would send the gathers of k_spmv (and the host
// transpose) out of bounds
__global__ void k_check_cols(long long nnz, const int* __restrict__ ci, int n_cols, int* fail)
{
	const long long stride = (long long)gridDim.x * blockDim.x;
	bool bad = false;
	for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) bad |= ((unsigned)ci[k] >= (unsigned)n_cols);
	if (bad) *fail = 1;
}

__global__ void k_diag_cplxf(int n, const int* rp, const int* ci, const ZF* v, ZF* d)
{	// clcg_smCcsr_get_diagonal (lcg_complex_cuda.cu), single precision
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	ZF out(0.f, 0.f);
	for (int k = rp[i]; k < rp[i + 1]; k++) if (ci[k] == i) { out = v[k]; break; }
	d[i] = out;
}

template <class T>
void host_transpose(int n_rows, int n_cols, const int* rp, const int* ci, const T* v, std::vector<int>& trp, std::vector<int>& tci, std::vector<T>& tv)
{
	const int nnz = rp[n_rows];
	trp.assign((size_t)n_cols + 1, 0); tci.resize((size_t)nnz); tv.resize((size_t)nnz);
	for (int k = 0; k < nnz; k++) trp[(size_t)ci[k] + 1]++;
	for (int i = 0; i < n_cols; i++) trp[(size_t)i + 1] += trp[(size_t)i];
	std::vector<int> fill(trp.begin(), trp.end() - 1);
	for (int i = 0; i < n_rows; i++)
		for (int k = rp[i]; k < rp[i + 1]; k++) { int d = fill[(size_t)ci[k]]++; tci[(size_t)d] = i; tv[(size_t)d] = v[k]; }
}

constexpr int kPad = 16;   // elements of zero padding behind col/val so that 128-bit tile loads never run off the end

template <class T>
void upload_csr(int n_rows, int nnz, const int* rp_h, const int* ci, const T* v, bool src_device,
	int** d_rp, int** d_ci, void** d_v, int4** d_tiles, int* n_tiles, int tile_nnz, int* lpr_out, int* chunk_out)
{
	*d_rp = dev_alloc<int>((size_t)n_rows + 1 + kPad);
	*d_ci = dev_alloc<int>((size_t)nnz + kPad);
	T* dv = dev_alloc<T>((size_t)nnz + kPad);
	*d_v = dv;
	LCG_CUDA_CHECK(cudaMemset(*d_ci + nnz, 0, kPad * sizeof(int)));
	LCG_CUDA_CHECK(cudaMemset(dv + nnz, 0, kPad * sizeof(T)));
	LCG_CUDA_CHECK(cudaMemcpy(*d_rp, rp_h, ((size_t)n_rows + 1) * sizeof(int), cudaMemcpyHostToDevice));
	const cudaMemcpyKind kind = src_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	LCG_CUDA_CHECK(cudaMemcpy(*d_ci, ci, (size_t)nnz * sizeof(int), kind));
	LCG_CUDA_CHECK(cudaMemcpy(dv, v, (size_t)nnz * sizeof(T), kind));
	std::vector<int4> tiles;
	const int lpr = pick_lpr(nnz, n_rows);
	const int rows_cap = pick_tile_rows(nnz, n_rows, lpr, tile_nnz);
	build_tiles(rp_h, n_rows, tile_nnz, rows_cap, tiles);
	*lpr_out = lpr; *chunk_out = pick_chunk(rows_cap);
	*n_tiles = (int)tiles.size();
	*d_tiles = dev_alloc<int4>(tiles.size());
	LCG_CUDA_CHECK(cudaMemcpy(*d_tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice));
}

// one thread per row: value -> index in vdict (bitwise match), (col - row) -> index in odict; *fail = 1 if either is missing
__global__ void k_dict_encode(int n_rows, const int* __restrict__ rp, const int* __restrict__ ci, const double* __restrict__ v,
	const double* __restrict__ vdict, int nv, const int* __restrict__ odict, int no, unsigned short* __restrict__ code, int* fail)
{
	__shared__ long long s_v[256];
	__shared__ int s_o[256];
	if (threadIdx.x < 256) { s_v[threadIdx.x] = __double_as_longlong(vdict[threadIdx.x]); s_o[threadIdx.x] = odict[threadIdx.x]; }
	__syncthreads();
	const int row = blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= n_rows) return;
	for (int k = rp[row]; k < rp[row + 1]; k++)
	{
		const long long bits = __double_as_longlong(v[k]);
		const int off = ci[k] - row;
		int vi = -1, oi = -1;
		for (int i = 0; i < nv; i++) if (s_v[i] == bits) { vi = i; break; }
		for (int i = 0; i < no; i++) if (s_o[i] == off) { oi = i; break; }
		if (vi < 0 || oi < 0) { *fail = 1; return; }
		code[k] = (unsigned short)(vi | (oi << 8));
	}
}

// ---- row patterns (k_spmv_pat): rows whose sequences of 16-bit codes are identical share one pattern ----------------
constexpr int kPatTab = 2048;   // open-addressing table of row hashes (patterns are few: 27 for a 3-D stencil)

__device__ __forceinline__ unsigned long long row_hash(const int* rp, const unsigned short* code, int row)
{
	unsigned long long h = 1469598103934665603ull;   // FNV-1a over (length, codes)
	const int kb = rp[row], ke = rp[row + 1];
	h = (h ^ (unsigned long long)(ke - kb)) * 1099511628211ull;
	for (int k = kb; k < ke; k++) h = (h ^ (unsigned long long)code[k]) * 1099511628211ull;
	return h ? h : 1ull;
}

// every row inserts its hash; slot_of_row[row] = table slot, rep[slot] = smallest row with that hash; *fail on overflow
__global__ void k_pat_insert(int n_rows, const int* __restrict__ rp, const unsigned short* __restrict__ code, unsigned long long* key, int* rep,
	int* slot_of_row, int* fail)
{
	const int row = blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= n_rows) return;
	const unsigned long long h = row_hash(rp, code, row);
	int slot = (int)(h % kPatTab);
	for (int probe = 0; probe < kPatTab; probe++)
	{
		const unsigned long long prev = atomicCAS(&key[slot], 0ull, h);
		if (prev == 0ull || prev == h) { atomicMin(&rep[slot], row); slot_of_row[row] = slot; return; }
		slot = (slot + 1) % kPatTab;
	}
	*fail = 1;
}

// pat[row] = pattern id of its slot; a row that differs from its pattern's representative (hash collision) sets *fail
__global__ void k_pat_assign(int n_rows, const int* __restrict__ rp, const unsigned short* __restrict__ code, const int* __restrict__ slot_of_row,
	const int* __restrict__ id_of_slot, const int* __restrict__ rep, unsigned char* pat, int* fail)
{
	const int row = blockIdx.x * blockDim.x + threadIdx.x;
	if (row >= n_rows) return;
	const int slot = slot_of_row[row];
	const int r = rep[slot];
	const int kb = rp[row], len = rp[row + 1] - kb, kr = rp[r];
	if (rp[r + 1] - kr != len) { *fail = 1; return; }
	for (int j = 0; j < len; j++) if (code[kb + j] != code[kr + j]) { *fail = 1; return; }
	pat[row] = (unsigned char)id_of_slot[slot];
}

// pat_thread[32 it + lane] = the pattern id shared by the R rows of one thread of warp work item `it` of k_spmv_pat (rows
// (a R + q) S + 32 ib + lane, q < R; lanes with 32 ib + lane >= S own nothing), 255 when a row is missing (beyond n_rows)
// or the ids differ
__global__ void k_pat_items(long long n_rows, int S, int nib, int n_items, const unsigned char* __restrict__ pat, unsigned char* __restrict__ thread_code)
{
	const int it = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (it >= n_items) return;
	const int a = it / nib, ib = it - a * nib;
	const int i = ib * 32 + lane;
	int p0 = -1; bool bad = false;
	if (i < S)
		for (int q = 0; q < kPatRows; q++)
		{
			const long long row = ((long long)a * kPatRows + q) * S + i;
			const int pq = row < n_rows ? (int)pat[row] : -1;
			if (q == 0) p0 = pq;
			bad = bad || pq < 0 || pq != p0;
		}
	thread_code[(size_t)it * 32 + lane] = (unsigned char)((bad || p0 < 0) ? 255 : p0);
}

// rows per pattern
__global__ void k_pat_count(int n_rows, const unsigned char* __restrict__ pat, unsigned int* count)
{
	__shared__ unsigned int s_cnt[256];
	s_cnt[threadIdx.x] = 0u;
	__syncthreads();
	for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) atomicAdd(&s_cnt[pat[r]], 1u);
	__syncthreads();
	if (s_cnt[threadIdx.x]) atomicAdd(&count[threadIdx.x], s_cnt[threadIdx.x]);
}

// thread items of k_spmv_pat_box: thread (it, lane) owns the rows (a R + q) S + 64 ib + 2 lane + e, q < R, e < 2.  flags[32 it +
// lane] = what its rows lack of the box (csr.cuh: kBoxDrop*), 255 when they are no consistent sub-boxes: a row beyond the matrix
// or without row flags, "no left column" on the right row of the pair (or vice versa), "no lower line" anywhere but on the first
// line, "no upper line" anywhere but on the last, planes that differ between rows.  *n_odd counts the warps that hold such a thread.
__global__ void k_pat_box_flags(long long n_rows, int S, int nib, int n_items, const unsigned char* __restrict__ pat,
	const unsigned char* __restrict__ rowflags, unsigned char* __restrict__ flags, unsigned int* n_odd)
{
	__shared__ unsigned char s_rf[256];
	s_rf[threadIdx.x] = rowflags[threadIdx.x];
	__syncthreads();
	const int it = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (it >= n_items) return;
	const int a = it / nib, ib = it - a * nib;
	const int i = ib * 64 + 2 * lane;
	int out = 0; bool bad = false, any_row = false;
	int planes = -1;
	if (i >= S) bad = true;
	else for (int q = 0; q < kPatRows; q++)
		for (int e = 0; e < 2; e++)
		{
			const long long row = ((long long)a * kPatRows + q) * S + i + e;
			if (row >= n_rows) { bad = true; continue; }
			any_row = true;
			const int f = (int)s_rf[pat[row]];
			if (f & 0x80) { bad = true; continue; }
			if ((f & kBoxDropL) && e != 0) bad = true;
			if ((f & kBoxDropR) && e != 1) bad = true;
			if ((f & kBoxDropLow) && q != 0) bad = true;
			if ((f & kBoxDropHigh) && q != kPatRows - 1) bad = true;
			if (planes < 0) planes = f & ~15;
			if ((f & ~15) != planes) bad = true;
			out |= f;
		}
	// a slice is dropped for the whole thread, so the rows it is dropped FROM must be all the rows it would touch: the left value
	// serves every left row, the first line only row 0, the last line only row R - 1 — but both rows of the pair on that line
	if (!bad)
		for (int q = 0; q < kPatRows && !bad; q++)
			for (int e = 0; e < 2; e++)
			{
				const int f = (int)s_rf[pat[((long long)a * kPatRows + q) * S + i + e]];
				if ((out & kBoxDropL) && e == 0 && !(f & kBoxDropL)) bad = true;
				if ((out & kBoxDropR) && e == 1 && !(f & kBoxDropR)) bad = true;
				if ((out & kBoxDropLow) && q == 0 && !(f & kBoxDropLow)) bad = true;
				if ((out & kBoxDropHigh) && q == kPatRows - 1 && !(f & kBoxDropHigh)) bad = true;
			}
	flags[(size_t)it * 32 + lane] = (unsigned char)(bad ? 255 : out);
	// a thread that falls back costs its whole WARP the time of the chain walk: count warps
	const bool warp_odd = __any_sync(0xffffffffu, bad && any_row);
	if (lane == 0 && warp_odd) atomicAdd(n_odd, 1u);
}

// block items of k_spmv_pat_march, one block of 8 warps per item: bitem[16 bi + w] = the pattern all rows of warp w's item
// share (255 = mixed, 254 = the item has no rows), bitem[16 bi + 8] = 1 when every existing row of the block item is a
// subset of the geometry pattern gpat (sup == gpat with a non-empty mask)
__global__ void k_pat_bitems(long long n_rows, int S, int wx, int wy, int nibb, int gpat, const unsigned char* __restrict__ pat,
	const PatInfo* __restrict__ info, unsigned char* __restrict__ bitem)
{
	const long long bi = blockIdx.x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const long long ab = bi / nibb; const int ibb = (int)(bi - ab * nibb);
	const int wq = warp / wx, wi = warp % wx;
	const long long row0 = ((ab * wy + wq) * kPatRows) * S + (long long)(ibb * wx + wi) * 32 + lane;
	int p0 = -2; bool mixed = false, geo_bad = false, any_row = false;
	for (int q = 0; q < kPatRows; q++)
	{
		const long long row = row0 + (long long)q * S;
		const int pq = row < n_rows ? (int)pat[row] : -1;
		if (q == 0) p0 = pq;
		mixed = mixed || pq != p0 || pq < 0;
		if (pq >= 0) { any_row = true; geo_bad = geo_bad || info[pq].sup != gpat || info[pq].mask == 0ull; }
	}
	const int first = __shfl_sync(0xffffffffu, p0, 0);
	mixed = mixed || p0 != first;
	const bool w_mixed = __any_sync(0xffffffffu, mixed), w_rows = __any_sync(0xffffffffu, any_row);
	const int bad_block = __syncthreads_or(geo_bad ? 1 : 0);
	if (lane == 0) bitem[16 * bi + warp] = (unsigned char)(!w_rows ? 254 : (w_mixed ? 255 : first));
	if (threadIdx.x == 0) bitem[16 * bi + 8] = (unsigned char)(bad_block ? 0 : 1);
}

// Row-pattern copy on top of the dictionary codes.  Leaves the handle without it when there are more than 255 distinct
// rows, a row longer than 64 entries, or a chain table larger than shared memory.
void try_patterns(CsrHandle* h, const std::vector<int>& rp_h, const std::vector<double>& vd, const std::vector<int>& od)
{
	const int n = h->n_rows;
	unsigned long long* d_key = dev_alloc<unsigned long long>(kPatTab);
	int* d_rep = dev_alloc<int>(kPatTab); int* d_slot = dev_alloc<int>((size_t)n); int* d_fail = dev_alloc<int>(1);
	int* d_id = dev_alloc<int>(kPatTab); unsigned char* d_pat = dev_alloc<unsigned char>((size_t)n);
	unsigned char* d_tab = nullptr; unsigned char* d_thread = nullptr; unsigned char* d_bitem = nullptr; int4* d_segs = nullptr;
	PatMarch* march = nullptr;
	PatBox* box = nullptr; unsigned char* d_box_flags = nullptr;
	bool ok = false;
	try
	{
		std::vector<int> rep_init(kPatTab, 0x7fffffff);
		LCG_CUDA_CHECK(cudaMemset(d_key, 0, sizeof(unsigned long long) * kPatTab));
		LCG_CUDA_CHECK(cudaMemcpy(d_rep, rep_init.data(), sizeof(int) * kPatTab, cudaMemcpyHostToDevice));
		LCG_CUDA_CHECK(cudaMemset(d_fail, 0, sizeof(int)));
		k_pat_insert<<<(n + 255) / 256, 256>>>(n, h->row_ptr, h->code, d_key, d_rep, d_slot, d_fail);
		LCG_CUDA_CHECK(cudaGetLastError());
		int fail = 1;
		LCG_CUDA_CHECK(cudaMemcpy(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost));
		std::vector<int> rep(kPatTab), id(kPatTab, -1);
		LCG_CUDA_CHECK(cudaMemcpy(rep.data(), d_rep, sizeof(int) * kPatTab, cudaMemcpyDeviceToHost));
		std::vector<int> reps;   // representative row of each pattern, in slot order
		if (!fail)
		{
			for (int s = 0; s < kPatTab; s++) if (rep[(size_t)s] != 0x7fffffff) { id[(size_t)s] = (int)reps.size(); reps.push_back(rep[(size_t)s]); }
			int maxlen = 0;
			for (int r : reps) maxlen = std::max(maxlen, rp_h[(size_t)r + 1] - rp_h[(size_t)r]);
			if (reps.size() <= 253 && maxlen >= 1 && maxlen <= 64)
			{
				LCG_CUDA_CHECK(cudaMemcpy(d_id, id.data(), sizeof(int) * kPatTab, cudaMemcpyHostToDevice));
				k_pat_assign<<<(n + 255) / 256, 256>>>(n, h->row_ptr, h->code, d_slot, d_id, d_rep, d_pat, d_fail);
				LCG_CUDA_CHECK(cudaGetLastError());
				LCG_CUDA_CHECK(cudaMemcpy(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost));
				if (!fail)
				{	// decode each representative row through the dictionaries, pick the stride on the longest one, build chains, plan, masks
					const size_t np = reps.size();
					std::vector<std::vector<std::pair<int, double>>> rows(np);
					std::vector<unsigned short> cbuf((size_t)maxlen);
					// the geometry pattern: the longest row, the most frequent one among equally long rows (the interior row of a
					// stencil; on a row block the ghost-coupled rows of the first and last plane are as long but fewer)
					std::vector<unsigned int> count(256, 0u);
					{
						unsigned int* d_count = reinterpret_cast<unsigned int*>(d_key);   // the hash table is not needed any more
						LCG_CUDA_CHECK(cudaMemset(d_count, 0, 256 * sizeof(unsigned int)));
						k_pat_count<<<592, 256>>>(n, d_pat, d_count);
						LCG_CUDA_CHECK(cudaGetLastError());
						LCG_CUDA_CHECK(cudaMemcpy(count.data(), d_count, 256 * sizeof(unsigned int), cudaMemcpyDeviceToHost));
					}
					size_t longest = 0;
					for (size_t p = 0; p < np; p++)
					{
						const int r = reps[p], kb = rp_h[(size_t)r], len = rp_h[(size_t)r + 1] - kb;
						if (len > 0) LCG_CUDA_CHECK(cudaMemcpy(cbuf.data(), h->code + kb, sizeof(unsigned short) * (size_t)len, cudaMemcpyDeviceToHost));
						for (int j = 0; j < len; j++) rows[p].push_back({od[cbuf[(size_t)j] >> 8], vd[cbuf[(size_t)j] & 255u]});
						if (rows[p].size() > rows[longest].size() || (rows[p].size() == rows[longest].size() && count[p] > count[longest])) longest = p;
					}
					static const char* s_env = getenv("LCGB200_PAT_STRIDE");   // experiments: force the stride (>= 32)
					int S = pat_pick_stride(rows[longest], n, kPatRows, kPatDefaultStride);
					if (s_env && atoi(s_env) >= 32) S = atoi(s_env);
					std::vector<std::vector<PatChainH>> chains(np);
					std::vector<PatInfo> info(np);
					std::vector<int> t0s(np, -1);
					size_t maxch = 1;
					for (size_t p = 0; p < np; p++)
					{
						pat_build_chains(rows[p], S, chains[p], &t0s[p]);
						info[p].info = (int)chains[p].size() | ((t0s[p] + 1) << 8);
						maxch = std::max(maxch, chains[p].size());
					}
					// 32-bit row and column arithmetic in the kernels: the last (partly empty) item and its reads stay below 2^31
					const long long n_super = ((long long)n + S - 1) / S;
					const long long n_a = (n_super + kPatRows - 1) / kPatRows, nib = (S + 31) / 32;
					const bool fits32 = (long long)n + (long long)h->n_cols + (long long)(4 * kPatRows + 4) * S < 0x7fffffffLL - 64 && n_a * nib < 0x7fffffffLL / 32;
					// the plane-marching kernel is opt-in (LCGB200_PAT_MARCH=1, read per handle): measured slower than the plain-load
					// kernel on one B200 (profiles/README_r02.md), kept for systems whose x does not stay in L2
					PatMarchH plan;
					const char* m_env = getenv("LCGB200_PAT_MARCH");
					if (fits32 && m_env && atoi(m_env) > 0) pat_plan_march(chains[longest], t0s[longest], S, n, h->n_cols, kPatRows, plan);   // reorders the geometry pattern's chains
					{
						std::vector<int> sup; std::vector<unsigned long long> mask;
						pat_build_masks(rows, chains, S, sup, mask);
						for (size_t p = 0; p < np; p++) { info[p].sup = sup[p]; info[p].mask = mask[p]; }
					}
					if (fits32 && np * maxch <= (size_t)kPatMaxChains)
					{
						// one table: chains | info (the order the kernels expect)
						std::vector<unsigned char> tab(np * maxch * sizeof(PatChainH) + np * sizeof(PatInfo), 0);
						PatChainH* tc = reinterpret_cast<PatChainH*>(tab.data());
						PatInfo* ti = reinterpret_cast<PatInfo*>(tab.data() + np * maxch * sizeof(PatChainH));
						for (size_t p = 0; p < np; p++) { std::copy(chains[p].begin(), chains[p].end(), tc + p * maxch); ti[p] = info[p]; }
						const int n_items = (int)(n_a * nib);
						d_tab = dev_alloc<unsigned char>(tab.size()); d_thread = dev_alloc<unsigned char>((size_t)n_items * 32);
						LCG_CUDA_CHECK(cudaMemcpy(d_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice));
						k_pat_items<<<(unsigned)(((long long)n_items * 32 + 255) / 256), 256>>>(n, S, (int)nib, n_items, d_pat, d_thread);
						LCG_CUDA_CHECK(cudaGetLastError());
						if (plan.ok)
						{
							const long long n_ab = (n_a + plan.wy - 1) / plan.wy;
							const int nibb = S / (32 * plan.wx);
							const long long n_bi = n_ab * nibb;
							// segments: 4 per resident block (2 blocks per SM; measured: 2 -> 0.229 ms, 4 -> 0.173 ms, 8 -> 0.183 ms at 27-point 256^3), at least 6 items each
							int sms = 148;
							{ int dev = 0, v = 0; if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v; }
							static const char* l_env = getenv("LCGB200_PAT_SEGS");   // experiments: segments per resident block
							const int per_block = (l_env && atoi(l_env) > 0) ? atoi(l_env) : 4;
							std::vector<PatSegH> segs;
							pat_build_segments(plan, n, S, kPatRows, per_block * 2 * sms, 6, segs);
							if (!segs.empty() && n_bi < 0x7fffffffLL / 16)
							{
								d_bitem = dev_alloc<unsigned char>((size_t)n_bi * 16); d_segs = dev_alloc<int4>(segs.size());
								LCG_CUDA_CHECK(cudaMemset(d_bitem, 0, (size_t)n_bi * 16));
								LCG_CUDA_CHECK(cudaMemcpy(d_segs, segs.data(), segs.size() * sizeof(int4), cudaMemcpyHostToDevice));
								const PatInfo* d_info = reinterpret_cast<const PatInfo*>(d_tab + np * maxch * sizeof(PatChainH));
								k_pat_bitems<<<(unsigned)n_bi, 256>>>(n, S, plan.wx, plan.wy, nibb, (int)longest, d_pat, d_info, d_bitem);
								LCG_CUDA_CHECK(cudaGetLastError());
								march = new PatMarch();
								march->gpat = (int)longest; march->G = plan.G; march->S2 = plan.S2; march->o0 = plan.o0; march->nlines = plan.nlines;
								march->wx = plan.wx; march->wy = plan.wy; march->dAb = plan.dAb; march->n_segs = (int)segs.size();
								for (int g = 0; g < kPatMaxPlanes; g++)
								{
									march->gbegin |= (unsigned int)(plan.group_begin[g] & 255) << (8 * g);
									march->gplane |= (unsigned int)(plan.group_plane[g] & 3) << (2 * g);
								}
								for (int g = plan.G; g < kPatMaxPlanes; g++) march->gbegin |= (unsigned int)(plan.group_begin[plan.G] & 255) << (8 * g);   // group_begin(G) for G < 4
								march->gend4 = (unsigned int)plan.group_begin[plan.G];
								// windows loaded ahead of the G an item reads: as many as two blocks per SM leave room for (at least one)
								static const char* a_env = getenv("LCGB200_PAT_AHEAD");
								int ahead = (a_env && atoi(a_env) > 0) ? atoi(a_env) : 3;
								while (ahead > 1 && pat_march_smem_bytes(plan.G + ahead, plan.nlines, plan.wx, (int)np, (int)maxch) > (size_t)kPatMarchSmemFor2) ahead--;
								march->nst = plan.G + ahead;
							}
						}
						// the box kernel: dense box stencils on grids whose lines and planes align with the threads' columns
						static const bool no_box = getenv("LCGB200_PAT_NO_BOX") != nullptr;   // comparison runs
						PatBoxH bplan; std::vector<unsigned char> rowflags;
						if (!no_box && !plan.ok && S % 2 == 0)
						{
							std::vector<int> sup2(np); std::vector<unsigned long long> mask2(np);
							for (size_t p = 0; p < np; p++) { sup2[p] = info[p].sup; mask2[p] = info[p].mask; }
							pat_plan_box(chains[longest], S, (int)longest, sup2, mask2, bplan, rowflags);
						}
						if (bplan.ok)
						{
							const int nib64 = (S + 63) / 64;
							const long long n_items64 = n_a * nib64;
							rowflags.resize(256, 0x80);
							unsigned char* d_rf = dev_alloc<unsigned char>(256);
							unsigned int* d_odd = reinterpret_cast<unsigned int*>(d_key);
							d_box_flags = dev_alloc<unsigned char>((size_t)n_items64 * 32);
							LCG_CUDA_CHECK(cudaMemcpy(d_rf, rowflags.data(), 256, cudaMemcpyHostToDevice));
							LCG_CUDA_CHECK(cudaMemset(d_odd, 0, sizeof(unsigned int)));
							k_pat_box_flags<<<(unsigned)((n_items64 * 32 + 255) / 256), 256>>>(n, S, nib64, (int)n_items64, d_pat, d_rf, d_box_flags, d_odd);
							LCG_CUDA_CHECK(cudaGetLastError());
							unsigned int n_odd = 0;
							LCG_CUDA_CHECK(cudaMemcpy(&n_odd, d_odd, sizeof(unsigned int), cudaMemcpyDeviceToHost));
							cudaFree(d_rf);
							// worth it only when few warps hold a thread that falls back to the chain tables
							if ((double)n_odd <= 0.10 * (double)n_items64)
							{
								box = new PatBox();
								box->G = bplan.G;
								for (int g = 0; g < bplan.G; g++)
								{
									box->center[g] = bplan.center[g];
									for (int dx = 0; dx < 3; dx++) for (int j = 0; j < 3; j++) box->coef[g][dx][j] = bplan.coef[g][dx][j];
								}
							}
							else { cudaFree(d_box_flags); d_box_flags = nullptr; }
						}
						LCG_CUDA_CHECK(cudaDeviceSynchronize());
						h->pat_box = box; h->pat_box_flags = d_box_flags;
						h->pat = d_pat; h->pat_thread = d_thread; h->pat_chain = d_tab; h->n_pat = (int)np;
						h->pat_maxch = (int)maxch; h->pat_stride = S; h->pat_nib = (int)nib; h->pat_items = n_items;
						h->pat_bitem = d_bitem; h->pat_segs = d_segs; h->pat_march = march;
						ok = true;
					}
				}
			}
		}
	}
	catch (...)
	{
		cudaFree(d_key); cudaFree(d_rep); cudaFree(d_slot); cudaFree(d_fail); cudaFree(d_id); cudaFree(d_pat); cudaFree(d_tab); cudaFree(d_thread);
		cudaFree(d_bitem); cudaFree(d_segs); delete march; cudaFree(d_box_flags); delete box;
		throw;
	}
	cudaFree(d_key); cudaFree(d_rep); cudaFree(d_slot); cudaFree(d_fail); cudaFree(d_id);
	if (!ok) { cudaFree(d_pat); cudaFree(d_tab); cudaFree(d_thread); cudaFree(d_bitem); cudaFree(d_segs); delete march; cudaFree(d_box_flags); delete box; }
}

// LCGB200_CSR_COMPRESS: if the matrix has <= 256 distinct values and <= 256 distinct (col - row) offsets, store a second
// copy as 16-bit codes + dictionaries (csr.cuh: k_spmv_dict).  The dictionaries are guessed from three windows of the
// matrix (start, middle, end) and then VERIFIED over every entry on the device; anything that does not fit leaves the
// handle uncompressed.  Returns true when the compressed copy exists.
bool try_compress(CsrHandle* h, const std::vector<int>& rp_h)
{
	const int n = h->n_rows, nnz = h->nnz;
	if (nnz <= 0) return false;
	for (int i = 0; i < n; i++) if (rp_h[(size_t)i + 1] - rp_h[(size_t)i] > kDictTileNnz - 8) return false;   // no over-long-row path in the dict kernel
	const int W = 1 << 17;
	std::vector<long long> vset; std::vector<int> oset;
	std::vector<int> cbuf((size_t)std::min(W, nnz)); std::vector<double> vbuf(cbuf.size());
	const int starts[3] = {0, std::max(0, nnz / 2 - W / 2), std::max(0, nnz - W)};
	for (int w = 0; w < 3; w++)
	{
		const int k0 = starts[w], cnt = std::min(W, nnz - k0);
		LCG_CUDA_CHECK(cudaMemcpy(cbuf.data(), h->col + k0, sizeof(int) * (size_t)cnt, cudaMemcpyDeviceToHost));
		LCG_CUDA_CHECK(cudaMemcpy(vbuf.data(), (const double*)h->val + k0, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost));
		int row = (int)(std::upper_bound(rp_h.begin(), rp_h.end(), k0) - rp_h.begin()) - 1;
		for (int k = 0; k < cnt; k++)
		{
			while (rp_h[(size_t)row + 1] <= k0 + k) row++;
			long long bits; std::memcpy(&bits, &vbuf[(size_t)k], 8);
			if (std::find(vset.begin(), vset.end(), bits) == vset.end()) { vset.push_back(bits); if (vset.size() > 256) return false; }
			const int off = cbuf[(size_t)k] - row;
			if (std::find(oset.begin(), oset.end(), off) == oset.end()) { oset.push_back(off); if (oset.size() > 256) return false; }
		}
	}
	std::sort(oset.begin(), oset.end());
	std::vector<double> vd(256, 0.0); std::vector<int> od(256, 0);
	for (size_t i = 0; i < vset.size(); i++) std::memcpy(&vd[i], &vset[i], 8);
	for (size_t i = 0; i < oset.size(); i++) od[i] = oset[i];
	double* d_vd = dev_alloc<double>(256); int* d_od = dev_alloc<int>(256);
	unsigned short* d_code = dev_alloc<unsigned short>((size_t)nnz + kPad);
	int* d_fail = dev_alloc<int>(1);
	bool ok = false;
	try
	{
		LCG_CUDA_CHECK(cudaMemcpy(d_vd, vd.data(), 256 * sizeof(double), cudaMemcpyHostToDevice));
		LCG_CUDA_CHECK(cudaMemcpy(d_od, od.data(), 256 * sizeof(int), cudaMemcpyHostToDevice));
		LCG_CUDA_CHECK(cudaMemset(d_code + nnz, 0, kPad * sizeof(unsigned short)));
		LCG_CUDA_CHECK(cudaMemset(d_fail, 0, sizeof(int)));
		k_dict_encode<<<(n + 255) / 256, 256>>>(n, h->row_ptr, h->col, (const double*)h->val, d_vd, (int)vset.size(), d_od, (int)oset.size(), d_code, d_fail);
		LCG_CUDA_CHECK(cudaGetLastError());
		int fail = 1;
		LCG_CUDA_CHECK(cudaMemcpy(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost));
		ok = (fail == 0);
	}
	catch (...) { cudaFree(d_vd); cudaFree(d_od); cudaFree(d_code); cudaFree(d_fail); throw; }
	cudaFree(d_fail);
	if (!ok) { cudaFree(d_vd); cudaFree(d_od); cudaFree(d_code); return false; }
	// 2-byte entries make a whole pass of 256 rows fit one stage for rows of up to 28 entries: one THREAD per row (no
	// cross-lane butterfly, no per-lane predication, gathers of 32 consecutive rows coalesce into 2 lines); longer
	// rows get the fewest lanes per row whose pass still fits
	const double avg = (double)nnz / (double)n;
	int dlpr = 1;
	while (dlpr < 32 && avg * (kThreads / dlpr) > (double)kDictTileNnz) dlpr *= 2;
	h->dlpr = dlpr;
	std::vector<int4> tiles;
	const int rows_cap = pick_tile_rows(nnz, n, dlpr, kDictTileNnz);
	build_tiles(rp_h.data(), n, kDictTileNnz, rows_cap, tiles, 8);
	h->dtiles = dev_alloc<int4>(tiles.size());
	LCG_CUDA_CHECK(cudaMemcpy(h->dtiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice));
	h->n_dtiles = (int)tiles.size(); h->dchunk = pick_chunk(rows_cap);
	h->code = d_code; h->vdict = d_vd; h->odict = d_od; h->n_vdict = (int)vset.size(); h->n_odict = (int)oset.size();
	static const bool no_pat = getenv("LCGB200_NO_PATTERNS") != nullptr;   // comparison runs: dictionary level only
	if (!no_pat) try_patterns(h, rp_h, vd, od);
	return true;
}

// ---- IC(0) preconditioner (LCGB200_CSR_IC0) ---------------------------------------------------------------------------
template <class T> struct IcTraits;
template <> struct IcTraits<double> { typedef IcReal M; };
template <> struct IcTraits<double2> { typedef IcCplx M; };
template <> struct IcTraits<ZF> { typedef IcCplxF M; };

template <class V>
void upload_factor(IcDev& F, int n, const std::vector<int>& rp, const std::vector<int>& ci, const std::vector<V>& val, bool upper)
{
	std::vector<int> order;
	F.n = n;
	F.n_levels = level_order(n, rp.data(), ci.data(), upper, 32, order);
	F.n_pos = (int)order.size();
	F.rp = dev_alloc<int>(rp.size()); F.ci = dev_alloc<int>(ci.size()); F.val = dev_alloc<V>(val.size());
	F.order = dev_alloc<int>(order.size()); F.ready = dev_alloc<int>((size_t)n); F.ctl = dev_alloc<IcCtl>(1);
	LCG_CUDA_CHECK(cudaMemcpy(F.rp, rp.data(), rp.size() * sizeof(int), cudaMemcpyHostToDevice));
	LCG_CUDA_CHECK(cudaMemcpy(F.ci, ci.data(), ci.size() * sizeof(int), cudaMemcpyHostToDevice));
	LCG_CUDA_CHECK(cudaMemcpy(F.val, val.data(), val.size() * sizeof(V), cudaMemcpyHostToDevice));
	LCG_CUDA_CHECK(cudaMemcpy(F.order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
	LCG_CUDA_CHECK(cudaMemset(F.ready, 0, (size_t)n * sizeof(int)));
	LCG_CUDA_CHECK(cudaMemset(F.ctl, 0, sizeof(IcCtl)));
}

// factorise the lower triangle of the (symmetric) matrix on the host — the reference's sequential algorithm, ic0_host.h —
// and put L, U = L^T and their level orders on the device
template <class T>
void setup_ic0(CsrHandle* h, const std::vector<int>& rp_h, const int* col, const T* val, bool on_device)
{
	typedef typename IcTraits<T>::M M;
	typedef typename M::T V;
	static_assert(sizeof(V) == sizeof(T), "storage types are layout-compatible with the reference's");
	const int n = h->n_rows, nnz = h->nnz;
	if (h->n_cols != n) { set_error_msg("IC(0) needs a square (unpartitioned) operator"); throw ApiFailure{LCGB200_SIZE_NOT_MATCH}; }
	std::vector<int> ci_h((size_t)nnz); std::vector<V> v_h((size_t)nnz);
	const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
	if (nnz > 0)
	{
		LCG_CUDA_CHECK(cudaMemcpy(ci_h.data(), col, (size_t)nnz * sizeof(int), kind));
		LCG_CUDA_CHECK(cudaMemcpy(v_h.data(), val, (size_t)nnz * sizeof(V), kind));
	}
	std::vector<int> lrp((size_t)n + 1, 0), lci; std::vector<V> lv;
	for (int i = 0; i < n; i++)
	{
		for (int k = rp_h[(size_t)i]; k < rp_h[(size_t)i + 1]; k++) if (ci_h[(size_t)k] <= i) { lci.push_back(ci_h[(size_t)k]); lv.push_back(v_h[(size_t)k]); }
		lrp[(size_t)i + 1] = (int)lci.size();
	}
	if (!ic0_lower<M>(n, lrp.data(), lci.data(), lv.data()))
	{
		set_error_msg("IC(0): every row needs a diagonal entry and ascending column indices");
		throw ApiFailure{LCGB200_NULL_PRECONDITION_MATRIX};
	}
	// U = L^T by counting (rows of U: ascending columns, the diagonal first)
	std::vector<int> urp((size_t)n + 1, 0), uci(lci.size()); std::vector<V> uv(lv.size());
	for (size_t k = 0; k < lci.size(); k++) urp[(size_t)lci[k] + 1]++;
	for (int i = 0; i < n; i++) urp[(size_t)i + 1] += urp[(size_t)i];
	{
		std::vector<int> fill(urp.begin(), urp.end() - 1);
		for (int i = 0; i < n; i++)
			for (int k = lrp[(size_t)i]; k < lrp[(size_t)i + 1]; k++) { const int d = fill[(size_t)lci[(size_t)k]]++; uci[(size_t)d] = i; uv[(size_t)d] = lv[(size_t)k]; }
	}
	upload_factor<V>(h->icL, n, lrp, lci, lv, false);
	upload_factor<V>(h->icU, n, urp, uci, uv, true);
	h->ic_tmp = dev_alloc<V>((size_t)n);
	h->has_ic0 = true;
}

void free_factor(IcDev& F) { cudaFree(F.rp); cudaFree(F.ci); cudaFree(F.val); cudaFree(F.order); cudaFree(F.ready); cudaFree(F.ctl); F = IcDev(); }

template <class T>
void create_typed(CsrHandle* h, const int* row_ptr, const int* col, const T* val, int location, int tile_nnz)
{
	const bool dev = (location == LCGB200_DEVICE);
	std::vector<int> rp_h((size_t)h->n_rows + 1);
	if (dev) LCG_CUDA_CHECK(cudaMemcpy(rp_h.data(), row_ptr, rp_h.size() * sizeof(int), cudaMemcpyDeviceToHost));
	else std::memcpy(rp_h.data(), row_ptr, rp_h.size() * sizeof(int));
	if (rp_h[0] != 0 || rp_h[(size_t)h->n_rows] != h->nnz) { set_error_msg("row_ptr[0] must be 0 and row_ptr[n] must equal nnz"); throw ApiFailure{LCGB200_SIZE_NOT_MATCH}; }
	for (int i = 0; i < h->n_rows; i++) if (rp_h[(size_t)i + 1] < rp_h[(size_t)i]) { set_error_msg("row_ptr must be non-decreasing"); throw ApiFailure{LCGB200_SIZE_NOT_MATCH}; }
	upload_csr<T>(h->n_rows, h->nnz, rp_h.data(), col, val, dev, &h->row_ptr, &h->col, &h->val, &h->tiles, &h->n_tiles, tile_nnz, &h->lpr, &h->chunk);
	if (h->nnz > 0)
	{	// 0 <= col < n_cols, checked over every entry on the device copy before anything gathers through it
		int* d_fail = dev_alloc<int>(1);
		int fail = 1;
		cudaError_t e = cudaMemset(d_fail, 0, sizeof(int));
		if (e == cudaSuccess)
		{
			const int grid = (int)std::min<long long>(((long long)h->nnz + 255) / 256, 148 * 16);
			k_check_cols<<<grid, 256>>>(h->nnz, h->col, h->n_cols, d_fail);
			e = cudaMemcpy(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost);
		}
		cudaFree(d_fail);
		LCG_CUDA_CHECK(e);
		if (fail) { set_error_msg("column index outside [0, n_cols)"); throw ApiFailure{LCGB200_SIZE_NOT_MATCH}; }
	}
	if (h->flags & LCGB200_CSR_TRANSPOSE)
	{
		std::vector<int> ci_h; std::vector<T> v_h;
		const int* cip = col; const T* vp = val;
		if (dev)
		{
			ci_h.resize((size_t)h->nnz); v_h.resize((size_t)h->nnz);
			LCG_CUDA_CHECK(cudaMemcpy(ci_h.data(), col, (size_t)h->nnz * sizeof(int), cudaMemcpyDeviceToHost));
			LCG_CUDA_CHECK(cudaMemcpy(v_h.data(), val, (size_t)h->nnz * sizeof(T), cudaMemcpyDeviceToHost));
			cip = ci_h.data(); vp = v_h.data();
		}
		std::vector<int> trp, tci; std::vector<T> tv;
		host_transpose<T>(h->n_rows, h->n_cols, rp_h.data(), cip, vp, trp, tci, tv);
		upload_csr<T>(h->n_cols, h->nnz, trp.data(), tci.data(), tv.data(), false, &h->t_row_ptr, &h->t_col, &h->t_val, &h->t_tiles, &h->t_n_tiles, tile_nnz, &h->t_lpr, &h->t_chunk);
	}
	if ((h->flags & LCGB200_CSR_COMPRESS) && std::is_same<T, double>::value) try_compress(h, rp_h);
	if (h->flags & LCGB200_CSR_IC0) setup_ic0<T>(h, rp_h, col, val, dev);
	if (h->flags & LCGB200_CSR_JACOBI)
	{
		T* d = dev_alloc<T>((size_t)h->n_rows);
		h->diag = d;
		const int grid = (h->n_rows + 255) / 256;
		if (std::is_same<T, double>::value) k_diag_real<<<grid, 256>>>(h->n_rows, h->row_ptr, h->col, (const double*)h->val, (double*)d);
		else if (std::is_same<T, ZF>::value) k_diag_cplxf<<<grid, 256>>>(h->n_rows, h->row_ptr, h->col, (const ZF*)h->val, (ZF*)d);
		else k_diag_cplx<<<grid, 256>>>(h->n_rows, h->row_ptr, h->col, (const double2*)h->val, (double2*)d);
		LCG_CUDA_CHECK(cudaGetLastError());
		LCG_CUDA_CHECK(cudaDeviceSynchronize());
	}
}

void destroy_handle(CsrHandle* h)
{
	if (!h) return;
	cudaFree(h->row_ptr); cudaFree(h->col); cudaFree(h->val); cudaFree(h->tiles);
	cudaFree(h->t_row_ptr); cudaFree(h->t_col); cudaFree(h->t_val); cudaFree(h->t_tiles);
	cudaFree(h->code); cudaFree(h->vdict); cudaFree(h->odict); cudaFree(h->dtiles);
	cudaFree(h->pat); cudaFree(h->pat_thread); cudaFree(h->pat_chain); cudaFree(h->pat_bitem); cudaFree(h->pat_segs); delete h->pat_march; cudaFree(h->pat_box_flags); delete h->pat_box;
	free_factor(h->icL); free_factor(h->icU); cudaFree(h->ic_tmp);
	cudaFree(h->diag); cudaFree(h->ws);
	cudaFree(h->d_state); cudaFree(h->d_partials);
	if (h->h_state) cudaFreeHost(h->h_state);
	if (h->h_state2) cudaFreeHost(h->h_state2);
	for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
	if (h->own_stream) cudaStreamDestroy(h->own_stream);
	if (h->own_event) cudaEventDestroy(h->own_event);
	delete h;
}

// ---- cuSPARSE dense-vector descriptors for user callbacks: resolved lazily, never touched by the built-in path
struct CusparseShim {
	void* lib = nullptr;
	int (*create)(lcgb200_dnvec_t*, long long, void*, int) = nullptr;   // cusparseCreateDnVec(descr*, int64 size, void* values, cudaDataType)
	int (*destroy)(lcgb200_dnvec_t) = nullptr;
	bool load()
	{
		if (create) return true;
		const char* names[] = {"libcusparse.so.12", "libcusparse.so"};
		for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
		if (!lib) { set_error_msg("user Ax callbacks need libcusparse (cusparseCreateDnVec) and it could not be loaded"); return false; }
		create = (int (*)(lcgb200_dnvec_t*, long long, void*, int))dlsym(lib, "cusparseCreateDnVec");
		destroy = (int (*)(lcgb200_dnvec_t))dlsym(lib, "cusparseDestroyDnVec");
		return create && destroy;
	}
};
CusparseShim g_cusparse;

struct DescrCache {
	int n; int dtype;   // CUDA_R_64F = 1, CUDA_C_64F = 5
	std::vector<std::pair<const void*, lcgb200_dnvec_t>> items;
	lcgb200_dnvec_t get(const void* p)
	{
		for (auto& it : items) if (it.first == p) return it.second;
		lcgb200_dnvec_t d = nullptr;
		if (g_cusparse.create(&d, (long long)n, const_cast<void*>(p), dtype) != 0) { set_error_msg("cusparseCreateDnVec failed"); throw CudaFailure(); }
		items.emplace_back(p, d);
		return d;
	}
	~DescrCache() { for (auto& it : items) g_cusparse.destroy(it.second); }
};

// ---- parameter validation, in the order the reference performs it ------------------------------------
int check_real(int solver_id, int n, const lcgb200_para& p, const void* m, const void* B, const void* lo, const void* hi)
{
	switch (solver_id)
	{
		case LCGB200_BICGSTAB2:	// lcg.cpp:819-825
			if (n <= 0) return LCGB200_INVILAD_VARIABLE_SIZE;
			if (p.max_iterations < 0) return LCGB200_INVILAD_MAX_ITERATIONS;
			if (p.epsilon <= 0.0) return LCGB200_INVILAD_EPSILON;
			if (p.restart_epsilon <= 0.0 || p.epsilon >= 1.0) return LCGB200_INVILAD_RESTART_EPSILON;
			break;
		case LCGB200_PG:	// lcg.cpp:1062-1065
			if (n <= 0) return LCGB200_INVILAD_VARIABLE_SIZE;
			if (p.max_iterations < 0) return LCGB200_INVILAD_MAX_ITERATIONS;
			if (p.epsilon <= 0.0) return LCGB200_INVILAD_EPSILON;
			if (p.step <= 0.0 || p.epsilon >= 1.0) return LCGB200_INVALID_LAMBDA;
			break;
		case LCGB200_SPG:	// lcg.cpp:1232-1238
			if (n <= 0) return LCGB200_INVILAD_VARIABLE_SIZE;
			if (p.max_iterations < 0) return LCGB200_INVILAD_MAX_ITERATIONS;
			if (p.epsilon <= 0.0 || p.epsilon >= 1.0) return LCGB200_INVILAD_EPSILON;
			if (p.step <= 0.0) return LCGB200_INVALID_LAMBDA;
			if (p.sigma <= 0.0 || p.sigma >= 1.0) return LCGB200_INVALID_SIGMA;
			if (p.beta <= 0.0 || p.beta >= 1.0) return LCGB200_INVALID_BETA;
			if (p.maxi_m <= 0) return LCGB200_INVALID_MAXIM;
			break;
		default:	// lcg.cpp:150-152
			if (n <= 0) return LCGB200_INVILAD_VARIABLE_SIZE;
			if (p.max_iterations < 0) return LCGB200_INVILAD_MAX_ITERATIONS;
			if (p.epsilon <= 0.0 || p.epsilon >= 1.0) return LCGB200_INVILAD_EPSILON;
	}
	if (!m || !B) return LCGB200_INVALID_POINTER;
	if ((solver_id == LCGB200_PG || solver_id == LCGB200_SPG) && (!lo || !hi)) return LCGB200_INVALID_POINTER;
	return 0;
}

int check_cplx(int n, const lcgb200_cpara& p, const void* m, const void* B)
{	// clcg.cpp:84-89
	if (n <= 0) return LCGB200_INVILAD_VARIABLE_SIZE;
	if (p.max_iterations < 0) return LCGB200_INVILAD_MAX_ITERATIONS;
	if (p.epsilon <= 0.0 || p.epsilon >= 1.0) return LCGB200_INVILAD_EPSILON;
	if (!m || !B) return LCGB200_C_INVALID_POINTER;
	return 0;
}

const lcgb200_para kDefPara = {0, 1e-6, 0, 1e-6, 1.0, 0.95, 0.9, 10};	// defparam, util.h:153
const lcgb200_cpara kDefCPara = {0, 1e-6, 0};				// defparam2, util.h:278

// The legacy default stream cannot be captured into a CUDA graph.  A solve on the built-in operator (no user callback that
// might enqueue work on a stream of its own) that was handed that stream is moved to a private non-blocking stream of the
// handle, ordered after everything the caller has enqueued so far; the entry points return only when the solve is complete.
cudaStream_t solve_stream(CsrHandle* h, bool builtin_only, cudaStream_t stream)
{
	if (!h || !builtin_only || !(stream == nullptr || stream == cudaStreamLegacy)) return stream;
	if (!h->own_stream)
	{
		LCG_CUDA_CHECK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
		LCG_CUDA_CHECK(cudaEventCreateWithFlags(&h->own_event, cudaEventDisableTiming));
	}
	LCG_CUDA_CHECK(cudaEventRecord(h->own_event, stream));
	LCG_CUDA_CHECK(cudaStreamWaitEvent(h->own_stream, h->own_event, 0));
	return h->own_stream;
}

inline double now_ms()
{
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// One solve, real.  `h` may be null (callback operator).  m/B/lo/hi live on the host unless dev_vecs.
int do_solve_real(CsrHandle* h, Operator<double>& A, int solver_id, double* m, const double* B, const double* lo, const double* hi,
	const lcgb200_para& para, int n, int n_ext, long long n_global,
	const std::function<ProgressFn(const double* m_dev)>& make_pf, bool dev_vecs, cudaStream_t stream, lcgb200_info* info,
	double* const* ws_out = nullptr, int n_ws_out = 0)
{
	const double t0 = now_ms();
	stream = solve_stream(h, A.h && !A.apply && !A.precond && !(bool)make_pf(nullptr), stream);   // a progress callback may use the caller's stream
	Engine E(stream, h);
	E.n_local = (size_t)n;
	const bool constrained = (solver_id == LCGB200_PG || solver_id == LCGB200_SPG);
	// m can be used in place only if it is a device vector, 16-byte aligned, and needs no ghost tail
	const bool m_inplace = dev_vecs && n_ext == n && ((reinterpret_cast<uintptr_t>(m) & 15) == 0);
	const bool b_inplace = dev_vecs && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
	const bool box_inplace = dev_vecs && constrained && ((reinterpret_cast<uintptr_t>(lo) & 15) == 0) && ((reinterpret_cast<uintptr_t>(hi) & 15) == 0);
	const size_t vec_bytes = (((size_t)n_ext * sizeof(double)) + 255) & ~size_t(255);
	int nvec = real_vector_count(solver_id) + (m_inplace ? 0 : 1) + (b_inplace ? 0 : 1) + ((constrained && !box_inplace) ? 2 : 0);
	const bool exact = reference_order();
	if (exact && E.multi()) { set_error_msg("reference-order mode runs on one GPU (a serial sum has no partition)"); throw ApiFailure{LCGB200_UNKNOWN_ERROR}; }
	if (exact) nvec += kMaxRed;   // one term per element and reduction slot
	E.reserve(vec_bytes * (size_t)nvec);
	if (exact) { E.terms_stride = vec_bytes / sizeof(double); E.d_terms = E.alloc<double>(E.terms_stride * kMaxRed); E.allocs.clear(); }
	double* d_m = m; const double* d_B = B; const double* d_lo = lo; const double* d_hi = hi;
	const cudaMemcpyKind in_kind = dev_vecs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	// copies of B (read by the init step only) come first; everything behind them is touched every iteration and forms the
	// persisting-L2 window (Engine::l2_window_begin)
	if (!b_inplace)
	{
		double* t = E.alloc<double>((size_t)n_ext);
		LCG_CUDA_CHECK(cudaMemcpyAsync(t, B, (size_t)n * sizeof(double), in_kind, stream));
		d_B = t;
	}
	if (!constrained) E.l2_from = E.ws_off;
	E.l2_unit = vec_bytes;
	if (!m_inplace)
	{
		d_m = E.alloc<double>((size_t)n_ext);
		LCG_CUDA_CHECK(cudaMemcpyAsync(d_m, m, (size_t)n * sizeof(double), in_kind, stream));
	}
	if (constrained && !box_inplace)
	{
		double* t1 = E.alloc<double>((size_t)n_ext); double* t2 = E.alloc<double>((size_t)n_ext);
		LCG_CUDA_CHECK(cudaMemcpyAsync(t1, lo, (size_t)n * sizeof(double), in_kind, stream));
		LCG_CUDA_CHECK(cudaMemcpyAsync(t2, hi, (size_t)n * sizeof(double), in_kind, stream));
		d_lo = t1; d_hi = t2;
	}
	DevState init;
	std::memset(&init, 0, sizeof(init));
	init.eps = para.epsilon; init.restart_eps = para.restart_epsilon; init.sigma = para.sigma; init.ls_beta = para.beta;
	init.n_global = n_global; init.abs_diff = para.abs_diff; init.max_it = para.max_iterations;
	init.sc[SC_STEP] = para.step;
	init.ret = RC_UNKNOWN;
	E.pf = make_pf(d_m);
	// user operator callbacks (cuSPARSE-descriptor Ax / Mx or host callbacks): one loop head per host round trip, so that
	// Afp / Mfp are never invoked after the convergence test has fired (the reference's call pattern)
	E.sync_each = A.host_side || (bool)A.apply || (bool)A.precond;
	if (E.comm && E.comm->poisoned()) { set_error_msg("this communicator timed out in an earlier solve: its sequence counters may disagree with the peers'; create a new one"); throw ApiFailure{LCGB200_UNKNOWN_ERROR}; }
	E.start(init);
	const size_t first_work = E.allocs.size();
	int ret = exact ? solve_real_x(E, A, solver_id, d_m, d_B, d_lo, d_hi, para, (size_t)n, (size_t)n_ext)
	                : solve_real(E, A, solver_id, d_m, d_B, d_lo, d_hi, para, (size_t)n, (size_t)n_ext);
	const double dev_ms = E.device_ms();
	// lcg() / lcgs(): the caller's work vectors receive the solver's (host arrays, in the solver's allocation order)
	for (int i = 0; i < n_ws_out && first_work + (size_t)i < E.allocs.size(); i++)
		if (ws_out[i]) LCG_CUDA_CHECK(cudaMemcpy(ws_out[i], E.allocs[first_work + (size_t)i], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
	if (E.multi() && E.comm->check_abort()) set_error_msg("a cross-GPU wait timed out (a peer rank never arrived): the solve was ended and the communicator is poisoned");
	if (!m_inplace)
	{
		const cudaMemcpyKind out_kind = dev_vecs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
		LCG_CUDA_CHECK(cudaMemcpyAsync(m, d_m, (size_t)n * sizeof(double), out_kind, stream));
	}
	LCG_CUDA_CHECK(cudaStreamSynchronize(stream));
	if (info)
	{
		info->iterations = E.h_st->k_report; info->checks = E.h_st->checks; info->spmv_launches = E.spmv_launches;
		info->kernel_launches = E.launches; info->residual = E.h_st->residual; info->device_ms = dev_ms; info->total_ms = now_ms() - t0;
		double pms[2]; int pct[2]; E.prof_collect(pms, pct);
		info->spmv_ms = pms[0]; info->vec_ms = pms[1]; info->spmv_timed = pct[0]; info->vec_timed = pct[1];
	}
	return ret;
}

inline int run_complex(Engine& E, const Operator<double2>& A, int id, double2* m, const double2* B, const lcgb200_cpara& p, size_t n, size_t ne)
{
	return E.d_terms ? solve_complex_x(E, A, id, m, B, p, n, ne) : solve_complex(E, A, id, m, B, p, n, ne);
}
inline int run_complex(Engine& E, const Operator<ZF>& A, int id, ZF* m, const ZF* B, const lcgb200_cpara& p, size_t n, size_t ne) { return solve_complexf(E, A, id, m, B, p, n, ne); }

// ZV = double2 (cuDoubleComplex vectors) or ZF (cuComplex vectors, clcg_cudaf.h)
template <class T> struct Ident { typedef T type; };   // keeps make_pf out of template argument deduction
template <class ZV>
int do_solve_cplx(CsrHandle* h, Operator<ZV>& A, int solver_id, ZV* m, const ZV* B, const lcgb200_cpara& para,
	int n, int n_ext, long long n_global, const std::function<ProgressFn(const typename Ident<ZV>::type* m_dev)>& make_pf, bool dev_vecs,
	cudaStream_t stream, lcgb200_info* info)
{
	const double t0 = now_ms();
	stream = solve_stream(h, A.h && !A.apply && !A.precond && !(bool)make_pf(nullptr), stream);   // a progress callback may use the caller's stream
	Engine E(stream, h);
	E.n_local = (size_t)n;
	const bool m_inplace = dev_vecs && n_ext == n && ((reinterpret_cast<uintptr_t>(m) & 15) == 0);
	const bool b_inplace = dev_vecs && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
	const size_t vec_bytes = (((size_t)n_ext * sizeof(ZV)) + 255) & ~size_t(255);
	int nvec = complex_vector_count(solver_id) + (m_inplace ? 0 : 1) + (b_inplace ? 0 : 1);
	// reference-order mode: the double-precision loops only (the cuComplex entry points have no CPU reference to be identical to)
	const bool exact = reference_order() && std::is_same<ZV, double2>::value;
	if (exact && E.multi()) { set_error_msg("reference-order mode runs on one GPU (a serial sum has no partition)"); throw ApiFailure{LCGB200_UNKNOWN_ERROR}; }
	const size_t term_bytes = (((size_t)n_ext * sizeof(double)) + 255) & ~size_t(255);
	E.reserve(vec_bytes * (size_t)nvec + (exact ? term_bytes * kMaxRed : 0));
	if (exact) { E.terms_stride = term_bytes / sizeof(double); E.d_terms = E.alloc<double>(E.terms_stride * kMaxRed); E.allocs.clear(); }
	ZV* d_m = m; const ZV* d_B = B;
	const cudaMemcpyKind in_kind = dev_vecs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	if (!b_inplace) { ZV* t = E.alloc<ZV>((size_t)n_ext); LCG_CUDA_CHECK(cudaMemcpyAsync(t, B, (size_t)n * sizeof(ZV), in_kind, stream)); d_B = t; }
	E.l2_from = E.ws_off; E.l2_unit = vec_bytes;   // behind the copy of B: what every iteration touches (persisting-L2 window)
	if (!m_inplace) { d_m = E.alloc<ZV>((size_t)n_ext); LCG_CUDA_CHECK(cudaMemcpyAsync(d_m, m, (size_t)n * sizeof(ZV), in_kind, stream)); }
	DevState init;
	std::memset(&init, 0, sizeof(init));
	init.eps = para.epsilon; init.n_global = n_global; init.abs_diff = para.abs_diff; init.max_it = para.max_iterations;
	// the single-precision entry points exist only in the reference's CUDA library: its residual definition (clcg_cudaf.cu:142-176)
	init.cres_mode = std::is_same<ZV, ZF>::value ? 1 : settings().cres_mode;
	init.ret = RC_UNKNOWN;
	E.pf = make_pf(d_m);
	E.sync_each = A.host_side || (bool)A.apply || (bool)A.precond;
	if (E.comm && E.comm->poisoned()) { set_error_msg("this communicator timed out in an earlier solve: its sequence counters may disagree with the peers'; create a new one"); throw ApiFailure{LCGB200_UNKNOWN_ERROR}; }
	E.start(init);
	int ret = run_complex(E, A, solver_id, d_m, d_B, para, (size_t)n, (size_t)n_ext);
	const double dev_ms = E.device_ms();
	if (E.multi() && E.comm->check_abort()) set_error_msg("a cross-GPU wait timed out (a peer rank never arrived): the solve was ended and the communicator is poisoned");
	if (!m_inplace)
	{
		const cudaMemcpyKind out_kind = dev_vecs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
		LCG_CUDA_CHECK(cudaMemcpyAsync(m, d_m, (size_t)n * sizeof(ZV), out_kind, stream));
	}
	LCG_CUDA_CHECK(cudaStreamSynchronize(stream));
	if (info)
	{
		info->iterations = E.h_st->k_report; info->checks = E.h_st->checks; info->spmv_launches = E.spmv_launches;
		info->kernel_launches = E.launches; info->residual = E.h_st->residual; info->device_ms = dev_ms; info->total_ms = now_ms() - t0;
		double pms[2]; int pct[2]; E.prof_collect(pms, pct);
		info->spmv_ms = pms[0]; info->vec_ms = pms[1]; info->spmv_timed = pct[0]; info->vec_timed = pct[1];
	}
	return ret;
}

template <class F> int guarded(F&& f)
{
	try { return f(); }
	catch (const ApiFailure& a) { return a.code; }
	catch (const CudaFailure&) { return LCGB200_UNKNOWN_ERROR; }
	catch (const std::exception& e) { set_error_msg(e.what()); return LCGB200_UNKNOWN_ERROR; }
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

const char* lcgb200_last_error(void) { return lcgb200::last_error(); }
int lcgb200_version(void) { return 100; }
void lcgb200_set_shadow_seed(long seed) { settings().shadow_seed = seed; }
void lcgb200_set_complex_residual_mode(int mode) { settings().cres_mode = mode ? 1 : 0; }
void lcgb200_set_poll_interval(int it) { settings().poll = it > 0 ? it : 1; }
void lcgb200_set_profile(int on) { settings().profile = on ? 1 : 0; }
void lcgb200_set_fused_small(int on) { settings().fused_small = on ? 1 : 0; }
void lcgb200_set_spin_timeout_ms(long long ms) { settings().spin_timeout_ms = ms; }
void lcgb200_set_graphs(int mode) { settings().graphs = mode; }
void lcgb200_set_pdl(int mode) { settings().pdl = mode; }
void lcgb200_set_reference_order(int mode) { settings().reference_order = mode; }
void lcgb200_set_l2_persist(int mode) { settings().l2_persist = mode; }

// sentinels: recognised by address, never executed on the fused path
void lcgb200_csr_ax(void*, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t, lcgb200_dnvec_t, const int, const int) {}
void lcgb200_jacobi_mx(void*, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t, lcgb200_dnvec_t, const int, const int) {}
void lcgb200_ic0_mx(void*, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t, lcgb200_dnvec_t, const int, const int) {}
void lcgb200_ic0_cmx(void*, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t, lcgb200_dnvec_t, const int, const int, int) {}
void lcgb200_csr_cax(void*, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t, lcgb200_dnvec_t, const int, const int, int) {}
void lcgb200_jacobi_cmx(void*, lcgb200_cublas_t, lcgb200_cusparse_t, lcgb200_dnvec_t, lcgb200_dnvec_t, const int, const int, int) {}

int lcgb200_csr_create_rect(lcgb200_csr_t* out, int n_rows, int n_cols, int nnz, const int* row_ptr, const int* col,
	const void* val, int value_type, int location, unsigned flags)
{
	if (!out) return LCGB200_INVALID_POINTER;
	*out = nullptr;
	if (n_rows <= 0 || n_cols < n_rows || nnz < 0) return LCGB200_INVILAD_VARIABLE_SIZE;
	if (!row_ptr || (nnz > 0 && (!col || !val))) return LCGB200_INVALID_POINTER;
	CsrHandle* h = new CsrHandle();
	int rc = guarded([&]() {
		h->value_type = value_type; h->n_rows = n_rows; h->n_cols = n_cols; h->nnz = nnz; h->n_global = n_rows; h->flags = flags;
		LCG_CUDA_CHECK(cudaGetDevice(&h->device));
		if (value_type == LCGB200_REAL) create_typed<double>(h, row_ptr, col, (const double*)val, location, kTileNnzReal);
		else if (value_type == LCGB200_COMPLEX_FLOAT) create_typed<ZF>(h, row_ptr, col, (const ZF*)val, location, kTileNnzReal);
		else if (value_type == LCGB200_COMPLEX) create_typed<double2>(h, row_ptr, col, (const double2*)val, location, kTileNnzCplx);
		else throw ApiFailure{LCGB200_INVILAD_VARIABLE_SIZE};
		return 0;
	});
	if (rc != 0) { destroy_handle(h); return rc; }
	*out = reinterpret_cast<lcgb200_csr_t>(h);
	return 0;
}

int lcgb200_csr_create(lcgb200_csr_t* out, int n, int nnz, const int* row_ptr, const int* col, const void* val,
	int value_type, int location, unsigned flags)
{
	return lcgb200_csr_create_rect(out, n, n, nnz, row_ptr, col, val, value_type, location, flags);
}

int lcgb200_csr_destroy(lcgb200_csr_t A) { destroy_handle(reinterpret_cast<CsrHandle*>(A)); return 0; }

int lcgb200_csr_set_user(lcgb200_csr_t A, void* user)
{
	if (!A) return LCGB200_INVALID_POINTER;
	reinterpret_cast<CsrHandle*>(A)->user = user; return 0;
}

int lcgb200_csr_attach_transpose(lcgb200_csr_t A, lcgb200_csr_t At)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	CsrHandle* t = reinterpret_cast<CsrHandle*>(At);
	if (!h) return LCGB200_INVALID_POINTER;
	if (t && (t->value_type != h->value_type || t->n_rows != h->n_rows)) { set_error_msg("the transposed block must hold the same rows [r0, r1) and value type"); return LCGB200_SIZE_NOT_MATCH; }
	h->t_handle = t;
	return 0;
}

int lcgb200_csr_set_row_offset(lcgb200_csr_t A, long long first_global_row)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	if (!h || first_global_row < 0) return LCGB200_INVALID_POINTER;
	h->row_offset = first_global_row;
	return 0;
}

int lcgb200_csr_get_ic0(lcgb200_csr_t A, int* lnz, int* row_ptr_host, int* col_host, void* val_host, int* n_levels_lower, int* n_levels_upper)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	if (!h) return LCGB200_INVALID_POINTER;
	if (!h->has_ic0) return LCGB200_NULL_PRECONDITION_MATRIX;
	return guarded([&]() {
		int total = 0;
		LCG_CUDA_CHECK(cudaMemcpy(&total, h->icL.rp + h->n_rows, sizeof(int), cudaMemcpyDeviceToHost));
		if (lnz) *lnz = total;
		if (n_levels_lower) *n_levels_lower = h->icL.n_levels;
		if (n_levels_upper) *n_levels_upper = h->icU.n_levels;
		const size_t es = h->value_type == LCGB200_COMPLEX ? 16 : 8;
		if (row_ptr_host) LCG_CUDA_CHECK(cudaMemcpy(row_ptr_host, h->icL.rp, ((size_t)h->n_rows + 1) * sizeof(int), cudaMemcpyDeviceToHost));
		if (col_host) LCG_CUDA_CHECK(cudaMemcpy(col_host, h->icL.ci, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost));
		if (val_host) LCG_CUDA_CHECK(cudaMemcpy(val_host, h->icL.val, (size_t)total * es, cudaMemcpyDeviceToHost));
		return 0;
	});
}

// z = (L L^T)^-1 r on device vectors with the handle's IC(0) factor (the two triangular solves of one preconditioner application)
int lcgb200_csr_ic0_apply(lcgb200_csr_t A, const void* r_dev, void* z_dev, void* stream)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	if (!h || !r_dev || !z_dev) return LCGB200_INVALID_POINTER;
	if (!h->has_ic0) return LCGB200_NULL_PRECONDITION_MATRIX;
	cudaStream_t s = (cudaStream_t)stream;
	return guarded([&]() {
		Engine E(s, h);
		E.n_local = (size_t)h->n_rows;
		DevState init; std::memset(&init, 0, sizeof(init));
		E.start(init);
		if (h->value_type == LCGB200_REAL) E.ic0_solve<double>(h, (const double*)r_dev, (double*)z_dev);
		else if (h->value_type == LCGB200_COMPLEX) E.ic0_solve<double2>(h, (const double2*)r_dev, (double2*)z_dev);
		else E.ic0_solve<ZF>(h, (const ZF*)r_dev, (ZF*)z_dev);
		LCG_CUDA_CHECK(cudaGetLastError());
		return 0;
	});
}

// The factorisation alone, on the host (no GPU needed): lower triangle as CSR (row_ptr[n+1], ascending columns, diagonal last
// in every row), values factorised in place.  value_type LCGB200_REAL / LCGB200_COMPLEX / LCGB200_COMPLEX_FLOAT.
int lcgb200_ic0_factor_host(int n, const int* row_ptr, const int* col, void* val, int value_type)
{
	if (n <= 0) return LCGB200_INVILAD_VARIABLE_SIZE;
	if (!row_ptr || !col || !val) return LCGB200_INVALID_POINTER;
	bool ok = false;
	if (value_type == LCGB200_REAL) ok = ic0_lower<IcReal>(n, row_ptr, col, static_cast<double*>(val));
	else if (value_type == LCGB200_COMPLEX) ok = ic0_lower<IcCplx>(n, row_ptr, col, static_cast<cuDoubleComplex*>(val));
	else if (value_type == LCGB200_COMPLEX_FLOAT) ok = ic0_lower<IcCplxF>(n, row_ptr, col, static_cast<cuComplex*>(val));
	else return LCGB200_INVILAD_VARIABLE_SIZE;
	return ok ? 0 : LCGB200_NULL_PRECONDITION_MATRIX;
}

int lcgb200_csr_get_diagonal(lcgb200_csr_t A, void* diag_host)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	if (!h || !diag_host) return LCGB200_INVALID_POINTER;
	if (!h->diag) return LCGB200_NULL_PRECONDITION_MATRIX;
	const size_t es = h->value_type == LCGB200_COMPLEX ? sizeof(double2) : 8;   // double, or a pair of floats
	return guarded([&]() { LCG_CUDA_CHECK(cudaMemcpy(diag_host, h->diag, es * (size_t)h->n_rows, cudaMemcpyDeviceToHost)); return 0; });
}

int lcgb200_csr_info(lcgb200_csr_t A, int* n_rows, int* n_cols, int* nnz, int* n_tiles, int* lanes_per_row)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	if (!h) return LCGB200_INVALID_POINTER;
	if (n_rows) *n_rows = h->n_rows; if (n_cols) *n_cols = h->n_cols; if (nnz) *nnz = h->nnz;
	if (n_tiles) *n_tiles = h->n_tiles; if (lanes_per_row) *lanes_per_row = h->lpr;
	return 0;
}

int lcgb200_csr_pattern_kernel(lcgb200_csr_t A, int* kernel, int* stride, int* n_patterns)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	if (!h) return LCGB200_INVALID_POINTER;
	if (kernel) *kernel = !h->pat ? 0 : (h->pat_box ? 2 : (h->pat_march ? 3 : 1));
	if (stride) *stride = h->pat ? h->pat_stride : 0;
	if (n_patterns) *n_patterns = h->pat ? h->n_pat : 0;
	return 0;
}

int lcgb200_csr_format(lcgb200_csr_t A, int* compressed, int* n_values, int* n_offsets, long long* stream_bytes)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	if (!h) return LCGB200_INVALID_POINTER;
	const long long S = h->value_type == LCGB200_COMPLEX ? 16 : 8;
	if (compressed) *compressed = h->pat ? 2 : (h->code ? 1 : 0);
	if (n_values) *n_values = h->n_vdict;
	if (n_offsets) *n_offsets = h->n_odict;
	if (stream_bytes)
	{
		if (h->pat) *stream_bytes = (long long)h->n_rows + 2LL * h->n_rows * S;   // one id per row + x + y
		else *stream_bytes = (long long)h->nnz * (h->code ? 2 : (S + 4)) + ((long long)h->n_rows + 1) * 4 + 2LL * h->n_rows * S;
	}
	return 0;
}

long long lcgb200_csr_spmv_bytes(lcgb200_csr_t A)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	if (!h) return 0;
	const long long S = h->value_type == LCGB200_COMPLEX ? 16 : 8;
	return (long long)h->nnz * (S + 4) + ((long long)h->n_rows + 1) * 4 + 2LL * h->n_rows * S;
}

}  // extern "C"

// ---- stand-alone SpMV launchers (tests / bench / ncu) ----------------------------------------------------
namespace {

// dots[0] = w.y, dots[1] = y.y, dots[2] = x.y
struct EpiProbeReal {
	static constexpr int NRED = 3;
	static constexpr bool ACTIVE = true;
	const double* w; double* out;
	__device__ void begin(const DevState*) {}
	__device__ void row(int i, double yi, double xi, double* acc) const
	{
		acc[0] = fma(w ? w[i] : xi, yi, acc[0]); acc[1] = fma(yi, yi, acc[1]); acc[2] = fma(xi, yi, acc[2]);
	}
	__device__ void finish(DevState*, const double* tot) const { out[0] = tot[0]; out[1] = tot[1]; out[2] = tot[2]; }
};
struct EpiProbeCplx {
	static constexpr int NRED = 6;
	static constexpr bool ACTIVE = true;
	const double2* w; double* out;
	__device__ void begin(const DevState*) {}
	__device__ void row(int i, double2 yi, double2 xi, double* acc) const
	{
		double2 wi = w ? w[i] : xi;
		acc[0] += wi.x * yi.x + wi.y * yi.y; acc[1] += wi.x * yi.y - wi.y * yi.x;
		acc[2] += yi.x * yi.x + yi.y * yi.y;
		acc[4] += xi.x * yi.x + xi.y * yi.y; acc[5] += xi.x * yi.y - xi.y * yi.x;
	}
	__device__ void finish(DevState*, const double* tot) const { for (int i = 0; i < 6; i++) out[i] = tot[i]; }
};

// halo exchange in front of a stand-alone SpMV on a partitioned handle: push half only when k_spmv holds the receive half
void exchange_for_spmv(CsrHandle* h, const void* x, cudaStream_t s)
{
	const int eb = h->value_type == LCGB200_COMPLEX ? 16 : 8;
	if (h->halo_in_spmv()) h->comm->push(x, eb, s, h->d_state);
	else h->comm->halo(const_cast<void*>(x), eb, s, h->p2p_dev() != nullptr, h->d_state);
}

void ensure_state(CsrHandle* h, cudaStream_t s)
{
	Engine E(s, h);   // allocates the cached state block on first use
	E.n_local = (size_t)h->n_rows;
	DevState init; std::memset(&init, 0, sizeof(init));
	E.start(init);
}

}  // namespace

extern "C" {

int lcgb200_csr_spmv(lcgb200_csr_t A, const void* x, void* y, int op, void* stream)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	if (!h || !x || !y) return LCGB200_INVALID_POINTER;
	if (op != 0 && !h->t_row_ptr) return LCGB200_NULL_PRECONDITION_MATRIX;
	cudaStream_t s = (cudaStream_t)stream;
	return guarded([&]() {
		ensure_state(h, s);
		if (h->comm && h->comm->size() > 1 && op == 0) exchange_for_spmv(h, x, s);
		if (h->value_type == LCGB200_REAL)
		{
			if (op == 0) launch_spmv<double, false>(h->view<double>(), (const double*)x, (double*)y, EpiNone<double>{}, h->d_state, h->d_partials, s);
			else launch_spmv<double, false>(h->tview<double>(), (const double*)x, (double*)y, EpiNone<double>{}, h->d_state, h->d_partials, s);
		}
		else if (h->value_type == LCGB200_COMPLEX_FLOAT)
		{
			if (op == 0) launch_spmv<ZF, false>(h->view<ZF>(), (const ZF*)x, (ZF*)y, EpiNone<ZF>{}, h->d_state, h->d_partials, s);
			else if (op == 1) launch_spmv<ZF, false>(h->tview<ZF>(), (const ZF*)x, (ZF*)y, EpiNone<ZF>{}, h->d_state, h->d_partials, s);
			else launch_spmv<ZF, true>(h->tview<ZF>(), (const ZF*)x, (ZF*)y, EpiNone<ZF>{}, h->d_state, h->d_partials, s);
		}
		else
		{
			if (op == 0) launch_spmv<double2, false>(h->view<double2>(), (const double2*)x, (double2*)y, EpiNone<double2>{}, h->d_state, h->d_partials, s);
			else if (op == 1) launch_spmv<double2, false>(h->tview<double2>(), (const double2*)x, (double2*)y, EpiNone<double2>{}, h->d_state, h->d_partials, s);
			else launch_spmv<double2, true>(h->tview<double2>(), (const double2*)x, (double2*)y, EpiNone<double2>{}, h->d_state, h->d_partials, s);
		}
		LCG_CUDA_CHECK(cudaGetLastError());
		return 0;
	});
}

int lcgb200_csr_spmv_dot(lcgb200_csr_t A, const void* x, void* y, const void* w, double* dots_dev, void* stream)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(A);
	if (!h || !x || !y || !dots_dev) return LCGB200_INVALID_POINTER;
	cudaStream_t s = (cudaStream_t)stream;
	return guarded([&]() {
		ensure_state(h, s);
		if (h->comm && h->comm->size() > 1) exchange_for_spmv(h, x, s);
		if (h->value_type == LCGB200_REAL)
			launch_spmv<double, false>(h->view<double>(), (const double*)x, (double*)y, EpiProbeReal{(const double*)w, dots_dev}, h->d_state, h->d_partials, s);
		else if (h->value_type == LCGB200_COMPLEX_FLOAT) { set_error_msg("lcgb200_csr_spmv_dot: double and double-complex operators only"); throw ApiFailure{LCGB200_SIZE_NOT_MATCH}; }
		else
			launch_spmv<double2, false>(h->view<double2>(), (const double2*)x, (double2*)y, EpiProbeCplx{(const double2*)w, dots_dev}, h->d_state, h->d_partials, s);
		LCG_CUDA_CHECK(cudaGetLastError());
		return 0;
	});
}

// ------------------------------------------------------------------------------------------ handle-shaped
int lcgb200_solve(lcgb200_csr_t Ah, int solver_id, double* m, const double* B, const double* low, const double* hig,
	const lcgb200_para* param, lcgb200_progress_cuda_ptr Pfp, unsigned flags, void* stream, lcgb200_info* info)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(Ah);
	if (!h) return LCGB200_INVALID_POINTER;
	if (h->value_type != LCGB200_REAL) return LCGB200_SIZE_NOT_MATCH;
	const lcgb200_para para = param ? *param : kDefPara;
	if (solver_id < LCGB200_CG || solver_id > LCGB200_SPG) solver_id = LCGB200_CGS;	// lcg.cpp:76-78
	int rc = check_real(solver_id, h->n_rows, para, m, B, low, hig);
	if (rc) return rc;
	if (solver_id == LCGB200_PCG && !(((flags & LCGB200_USE_JACOBI) && h->diag) || ((flags & LCGB200_USE_IC0) && h->has_ic0))) return LCGB200_NULL_PRECONDITION_MATRIX;
	return guarded([&]() {
		Operator<double> A; A.h = h;
		if (solver_id == LCGB200_PCG) { if (flags & LCGB200_USE_IC0) A.ic0 = h; else A.diag = (const double*)h->diag; }
		auto make_pf = [&](const double* m_dev) -> ProgressFn {
			if (!Pfp) return ProgressFn();
			return [=](double res, int k) { return Pfp(h->user, m_dev, res, &para, h->n_rows, h->nnz, k); };
		};
		return do_solve_real(h, A, solver_id, m, B, low, hig, para, h->n_rows, h->n_cols, h->n_global, make_pf,
			(flags & LCGB200_VEC_DEVICE) != 0, (cudaStream_t)stream, info);
	});
}

int lcgb200_csolve(lcgb200_csr_t Ah, int solver_id, void* m, const void* B, const lcgb200_cpara* param,
	lcgb200_cprogress_cuda_ptr Pfp, unsigned flags, void* stream, lcgb200_info* info)
{
	CsrHandle* h = reinterpret_cast<CsrHandle*>(Ah);
	if (!h) return LCGB200_C_INVALID_POINTER;
	if (h->value_type != LCGB200_COMPLEX && h->value_type != LCGB200_COMPLEX_FLOAT) return LCGB200_C_SIZE_NOT_MATCH;
	const lcgb200_cpara para = param ? *param : kDefCPara;
	if (solver_id < LCGB200_CBICG || solver_id > LCGB200_CPCG) solver_id = LCGB200_CCGS;	// clcg.cpp:68-70
	int rc = check_cplx(h->n_rows, para, m, B);
	if (rc) return rc;
	if (solver_id == LCGB200_CBICG && !h->t_row_ptr && !h->t_handle)
	{
		set_error_msg("CLCG_BICG needs the transposed operator: create the handle with LCGB200_CSR_TRANSPOSE (a partitioned block: lcgb200_csr_attach_transpose)");
		return LCGB200_C_UNKNOWN_SOLVER;
	}
	if (solver_id == LCGB200_CPCG && !(((flags & LCGB200_USE_JACOBI) && h->diag) || ((flags & LCGB200_USE_IC0) && h->has_ic0))) return LCGB200_NULL_PRECONDITION_MATRIX;
	const bool use_ic0 = (flags & LCGB200_USE_IC0) != 0;
	return guarded([&]() {
		const int n_ext = std::max(h->n_cols, h->t_handle ? h->t_handle->n_cols : 0);
		if (h->value_type == LCGB200_COMPLEX_FLOAT)
		{	// cuComplex vectors (m, B: interleaved float pairs)
			Operator<ZF> A; A.h = h;
			if (solver_id == LCGB200_CPCG) { if (use_ic0) A.ic0 = h; else A.diag = (const ZF*)h->diag; }
			auto make_pf = [&](const ZF* m_dev) -> ProgressFn {
				if (!Pfp) return ProgressFn();
				return [=](double res, int k) { return Pfp(h->user, m_dev, res, &para, h->n_rows, h->nnz, k); };
			};
			return do_solve_cplx<ZF>(h, A, solver_id, (ZF*)m, (const ZF*)B, para, h->n_rows, n_ext, h->n_global, make_pf,
				(flags & LCGB200_VEC_DEVICE) != 0, (cudaStream_t)stream, info);
		}
		Operator<double2> A; A.h = h;
		if (solver_id == LCGB200_CPCG) { if (use_ic0) A.ic0 = h; else A.diag = (const double2*)h->diag; }
		auto make_pf = [&](const double2* m_dev) -> ProgressFn {
			if (!Pfp) return ProgressFn();
			return [=](double res, int k) { return Pfp(h->user, m_dev, res, &para, h->n_rows, h->nnz, k); };
		};
		return do_solve_cplx<double2>(h, A, solver_id, (double2*)m, (const double2*)B, para, h->n_rows, n_ext, h->n_global, make_pf,
			(flags & LCGB200_VEC_DEVICE) != 0, (cudaStream_t)stream, info);
	});
}

// --------------------------------------------------------------------------------------- reference-shaped
static int ref_real(lcgb200_axfunc_cuda_ptr Afp, lcgb200_axfunc_cuda_ptr Mfp, lcgb200_progress_cuda_ptr Pfp, double* m, const double* B,
	const double* low, const double* hig, const int n, const int nz, const lcgb200_para* param, void* instance,
	lcgb200_cublas_t cub, lcgb200_cusparse_t cus, int solver_id)
{
	const lcgb200_para para = param ? *param : kDefPara;
	int rc = check_real(solver_id, n, para, m, B, low, hig);
	if (rc) return rc;
	const bool builtin = (Afp == lcgb200_csr_ax);
	if (!builtin && (!cub || !cus)) return LCGB200_INVALID_POINTER;	// lcg_cuda.cu:97-98
	if (!Afp) return LCGB200_INVALID_POINTER;
	if (builtin)
	{
		CsrHandle* h = reinterpret_cast<CsrHandle*>(instance);
		if (!h) return LCGB200_INVALID_POINTER;
		if (h->value_type != LCGB200_REAL || h->n_rows != n) return LCGB200_SIZE_NOT_MATCH;
		if (solver_id == LCGB200_PCG)
		{
			if (Mfp == lcgb200_jacobi_mx) { if (!h->diag) return LCGB200_NULL_PRECONDITION_MATRIX; }
			else if (Mfp == lcgb200_ic0_mx) { if (!h->has_ic0) return LCGB200_NULL_PRECONDITION_MATRIX; }
			else if (!Mfp) return LCGB200_INVALID_POINTER;
		}
		return guarded([&]() {
			Operator<double> A; A.h = h;
			DescrCache dc{n, 1, {}};
			if (solver_id == LCGB200_PCG)
			{
				if (Mfp == lcgb200_jacobi_mx) A.diag = (const double*)h->diag;
				else if (Mfp == lcgb200_ic0_mx) A.ic0 = h;
				else
				{
					if (!g_cusparse.load()) throw ApiFailure{LCGB200_UNKNOWN_ERROR};
					A.precond = [&](const double* x, double* y, int) { Mfp(h->user, cub, cus, dc.get(x), dc.get(y), n, nz); };
				}
			}
			auto make_pf = [&](const double* m_dev) -> ProgressFn {
				if (!Pfp) return ProgressFn();
				return [=](double res, int k) { return Pfp(h->user, m_dev, res, &para, n, nz, k); };
			};
			return do_solve_real(h, A, solver_id, m, B, low, hig, para, n, h->n_cols, h->n_global, make_pf, false, nullptr, nullptr);
		});
	}
	if (solver_id == LCGB200_PCG && (!Mfp || Mfp == lcgb200_jacobi_mx)) return LCGB200_INVALID_POINTER;
	return guarded([&]() {
		if (!g_cusparse.load()) throw ApiFailure{LCGB200_UNKNOWN_ERROR};
		DescrCache dc{n, 1, {}};
		Operator<double> A;
		A.apply = [&](const double* x, double* y, int) { Afp(instance, cub, cus, dc.get(x), dc.get(y), n, nz); };
		if (solver_id == LCGB200_PCG) A.precond = [&](const double* x, double* y, int) { Mfp(instance, cub, cus, dc.get(x), dc.get(y), n, nz); };
		auto make_pf = [&](const double* m_dev) -> ProgressFn {
			if (!Pfp) return ProgressFn();
			return [=](double res, int k) { return Pfp(instance, m_dev, res, &para, n, nz, k); };
		};
		return do_solve_real(nullptr, A, solver_id, m, B, low, hig, para, n, n, n, make_pf, false, nullptr, nullptr);
	});
}

int lcgb200_solver_cuda(lcgb200_axfunc_cuda_ptr Afp, lcgb200_progress_cuda_ptr Pfp, double* m, const double* B,
	const int n_size, const int nz_size, const lcgb200_para* param, void* instance,
	lcgb200_cublas_t cub, lcgb200_cusparse_t cus, int solver_id)
{
	// lcg_cuda.cu:40-58: CG and CGS, anything else -> CG.  BICGSTAB / BICGSTAB2 are accepted as an extension.
	int id = LCGB200_CG;
	if (solver_id == LCGB200_CGS || solver_id == LCGB200_BICGSTAB || solver_id == LCGB200_BICGSTAB2) id = solver_id;
	return ref_real(Afp, nullptr, Pfp, m, B, nullptr, nullptr, n_size, nz_size, param, instance, cub, cus, id);
}

int lcgb200_solver_preconditioned_cuda(lcgb200_axfunc_cuda_ptr Afp, lcgb200_axfunc_cuda_ptr Mfp, lcgb200_progress_cuda_ptr Pfp,
	double* m, const double* B, const int n_size, const int nz_size, const lcgb200_para* param, void* instance,
	lcgb200_cublas_t cub, lcgb200_cusparse_t cus, int)
{
	return ref_real(Afp, Mfp, Pfp, m, B, nullptr, nullptr, n_size, nz_size, param, instance, cub, cus, LCGB200_PCG);
}

int lcgb200_solver_constrained_cuda(lcgb200_axfunc_cuda_ptr Afp, lcgb200_progress_cuda_ptr Pfp, double* m, const double* B,
	const double* low, const double* hig, const int n_size, const int nz_size, const lcgb200_para* param, void* instance,
	lcgb200_cublas_t cub, lcgb200_cusparse_t cus, int solver_id)
{
	// lcg.cpp:126-137: PG and SPG, anything else -> PG (the reference's CUDA build only has the broken lpg)
	const int id = (solver_id == LCGB200_SPG) ? LCGB200_SPG : LCGB200_PG;
	return ref_real(Afp, nullptr, Pfp, m, B, low, hig, n_size, nz_size, param, instance, cub, cus, id);
}

}  // extern "C"

// ZV = double2: clcg_cuda.h entry points; ZV = ZF: the cuComplex overloads of clcg_cudaf.h (progress callback with a float `converge`)
template <class ZV, class PfT>
static int ref_cplx(lcgb200_caxfunc_cuda_ptr Afp, lcgb200_caxfunc_cuda_ptr Mfp, PfT Pfp, void* m, const void* B,
	const int n, const int nz, const lcgb200_cpara* param, void* instance, lcgb200_cublas_t cub, lcgb200_cusparse_t cus, int solver_id)
{
	constexpr bool kF32 = std::is_same<ZV, ZF>::value;
	constexpr int kValueType = kF32 ? LCGB200_COMPLEX_FLOAT : LCGB200_COMPLEX, kCudaType = kF32 ? 4 : 5;   // CUDA_C_32F / CUDA_C_64F
	const lcgb200_cpara para = param ? *param : kDefCPara;
	int rc = check_cplx(n, para, m, B);
	if (rc) return rc;
	const bool builtin = (Afp == lcgb200_csr_cax);
	if (!builtin && (!cub || !cus)) return LCGB200_INVALID_POINTER;	// clcg_cuda.cu:100-101 (returns the LCG_ constant)
	if (!Afp) return LCGB200_INVALID_POINTER;
	if (builtin)
	{
		CsrHandle* h = reinterpret_cast<CsrHandle*>(instance);
		if (!h) return LCGB200_INVALID_POINTER;
		if (h->value_type != kValueType || h->n_rows != n) return LCGB200_C_SIZE_NOT_MATCH;
		if (solver_id == LCGB200_CBICG && !h->t_row_ptr && !h->t_handle) { set_error_msg("CLCG_BICG needs LCGB200_CSR_TRANSPOSE"); return LCGB200_C_UNKNOWN_SOLVER; }
		if (solver_id == LCGB200_CPCG)
		{
			if (Mfp == lcgb200_jacobi_cmx) { if (!h->diag) return LCGB200_NULL_PRECONDITION_MATRIX; }
			else if (Mfp == lcgb200_ic0_cmx) { if (!h->has_ic0) return LCGB200_NULL_PRECONDITION_MATRIX; }
			else if (!Mfp) return LCGB200_INVALID_POINTER;
		}
		return guarded([&]() {
			Operator<ZV> A; A.h = h;
			DescrCache dc{n, kCudaType, {}};
			if (solver_id == LCGB200_CPCG)
			{
				if (Mfp == lcgb200_jacobi_cmx) A.diag = (const ZV*)h->diag;
				else if (Mfp == lcgb200_ic0_cmx) A.ic0 = h;
				else
				{
					if (!g_cusparse.load()) throw ApiFailure{LCGB200_UNKNOWN_ERROR};
					A.precond = [&](const ZV* x, ZV* y, int) { Mfp(h->user, cub, cus, dc.get(x), dc.get(y), n, nz, 0); };
				}
			}
			auto make_pf = [&](const ZV* m_dev) -> ProgressFn {
				if (!Pfp) return ProgressFn();
				return [=](double res, int k) { return Pfp(h->user, m_dev, res, &para, n, nz, k); };
			};
			return do_solve_cplx<ZV>(h, A, solver_id, (ZV*)m, (const ZV*)B, para, n, std::max(h->n_cols, h->t_handle ? h->t_handle->n_cols : 0), h->n_global, make_pf, false, nullptr, nullptr);
		});
	}
	if (solver_id == LCGB200_CPCG && (!Mfp || Mfp == lcgb200_jacobi_cmx)) return LCGB200_INVALID_POINTER;
	return guarded([&]() {
		if (!g_cusparse.load()) throw ApiFailure{LCGB200_UNKNOWN_ERROR};
		DescrCache dc{n, kCudaType, {}};
		Operator<ZV> A;
		A.apply = [&](const ZV* x, ZV* y, int op) { Afp(instance, cub, cus, dc.get(x), dc.get(y), n, nz, op); };
		if (solver_id == LCGB200_CPCG) A.precond = [&](const ZV* x, ZV* y, int) { Mfp(instance, cub, cus, dc.get(x), dc.get(y), n, nz, 0); };
		auto make_pf = [&](const ZV* m_dev) -> ProgressFn {
			if (!Pfp) return ProgressFn();
			return [=](double res, int k) { return Pfp(instance, m_dev, res, &para, n, nz, k); };
		};
		return do_solve_cplx<ZV>(nullptr, A, solver_id, (ZV*)m, (const ZV*)B, para, n, n, n, make_pf, false, nullptr, nullptr);
	});
}

extern "C" {

int lcgb200_csolver_cuda(lcgb200_caxfunc_cuda_ptr Afp, lcgb200_cprogress_cuda_ptr Pfp, void* m, const void* B,
	const int n_size, const int nz_size, const lcgb200_cpara* param, void* instance,
	lcgb200_cublas_t cub, lcgb200_cusparse_t cus, int solver_id)
{
	// clcg_cuda.cu:42-60 knows BICG and BICG_SYM and returns CLCG_UNKNOWN_SOLVER otherwise; we add the three
	// solvers the reference only has on the CPU (clcg.cpp:366-882).
	switch (solver_id)
	{
		case LCGB200_CBICG: case LCGB200_CBICG_SYM: case LCGB200_CCGS: case LCGB200_CBICGSTAB: case LCGB200_CTFQMR: break;
		default: return LCGB200_C_UNKNOWN_SOLVER;
	}
	return ref_cplx<double2>(Afp, nullptr, Pfp, m, B, n_size, nz_size, param, instance, cub, cus, solver_id);
}

int lcgb200_csolver_preconditioned_cuda(lcgb200_caxfunc_cuda_ptr Afp, lcgb200_caxfunc_cuda_ptr Mfp, lcgb200_cprogress_cuda_ptr Pfp,
	void* m, const void* B, const int n_size, const int nz_size, const lcgb200_cpara* param, void* instance,
	lcgb200_cublas_t cub, lcgb200_cusparse_t cus, int solver_id)
{
	if (solver_id != LCGB200_CPCG) return LCGB200_C_UNKNOWN_SOLVER;	// clcg_cuda.cu:75-83
	return ref_cplx<double2>(Afp, Mfp, Pfp, m, B, n_size, nz_size, param, instance, cub, cus, LCGB200_CPCG);
}

// replace the cuComplex overloads of clcg_solver_cuda / clcg_solver_preconditioned_cuda (clcg_cudaf.h:81-83,103-105;
// clcg_cudaf.cu: BICG :86-252, BICG_SYM :254-401, PCG :403-558).  m, B: HOST arrays of n_size float pairs.
int lcgb200_csolver_cudaf(lcgb200_caxfunc_cuda_ptr Afp, lcgb200_cprogress_cudaf_ptr Pfp, void* m, const void* B,
	const int n_size, const int nz_size, const lcgb200_cpara* param, void* instance,
	lcgb200_cublas_t cub, lcgb200_cusparse_t cus, int solver_id)
{
	if (solver_id != LCGB200_CBICG && solver_id != LCGB200_CBICG_SYM) return LCGB200_C_UNKNOWN_SOLVER;	// clcg_cudaf.cu:42-60
	return ref_cplx<ZF>(Afp, nullptr, Pfp, m, B, n_size, nz_size, param, instance, cub, cus, solver_id);
}

int lcgb200_csolver_preconditioned_cudaf(lcgb200_caxfunc_cuda_ptr Afp, lcgb200_caxfunc_cuda_ptr Mfp, lcgb200_cprogress_cudaf_ptr Pfp,
	void* m, const void* B, const int n_size, const int nz_size, const lcgb200_cpara* param, void* instance,
	lcgb200_cublas_t cub, lcgb200_cusparse_t cus, int solver_id)
{
	if (solver_id != LCGB200_CPCG) return LCGB200_C_UNKNOWN_SOLVER;	// clcg_cudaf.cu:70-84
	return ref_cplx<ZF>(Afp, Mfp, Pfp, m, B, n_size, nz_size, param, instance, cub, cus, LCGB200_CPCG);
}

}  // extern "C"

// ================================================================================ CPU-shaped entry points
// lcg_solver / lcg_solver_preconditioned / lcg_solver_constrained (lcg.h:71-113) and clcg_solver (clcg.h:74-76): the
// reference's HOST-callback API.  With the sentinel callbacks (lcgb200_csr_ax_host & co.) and an lcgb200_csr_t as
// `instance` the solve runs on the fused built-in operator; any other callback is honoured on the generic path (the
// vector is staged to the host, the user's callback runs there, the product is staged back — correct, not fast).
// The progress callback receives a HOST copy of the current solution, as in the reference (lcg.h:53-54).
namespace {

template <class T>
struct HostStage {	// page-locked staging vectors for host callbacks
	T* x = nullptr; T* y = nullptr; T* m = nullptr; int n = 0;
	explicit HostStage(int n_) : n(n_) {}
	T* get(T*& p) { if (!p) LCG_CUDA_CHECK(cudaMallocHost((void**)&p, sizeof(T) * (size_t)n)); return p; }
	~HostStage() { if (x) cudaFreeHost(x); if (y) cudaFreeHost(y); if (m) cudaFreeHost(m); }
};

int host_real(lcgb200_axfunc_ptr Afp, lcgb200_axfunc_ptr Mfp, lcgb200_progress_ptr Pfp, double* m, const double* B, const double* low, const double* hig,
	const int n, const lcgb200_para* param, void* instance, int solver_id, double* const* ws_out = nullptr, int n_ws_out = 0)
{
	const lcgb200_para para = param ? *param : kDefPara;
	int rc = check_real(solver_id, n, para, m, B, low, hig);
	if (rc) return rc;
	if (!Afp) return LCGB200_INVALID_POINTER;
	const bool builtin = (Afp == lcgb200_csr_ax_host);
	CsrHandle* h = builtin ? reinterpret_cast<CsrHandle*>(instance) : nullptr;
	if (builtin)
	{
		if (!h) return LCGB200_INVALID_POINTER;
		if (h->value_type != LCGB200_REAL || h->n_rows != n) return LCGB200_SIZE_NOT_MATCH;
	}
	if (solver_id == LCGB200_PCG)
	{
		if (!Mfp) return LCGB200_INVALID_POINTER;
		if (Mfp == lcgb200_jacobi_mx_host && !(builtin && h->diag)) return LCGB200_NULL_PRECONDITION_MATRIX;
		if (Mfp == lcgb200_ic0_mx_host && !(builtin && h->has_ic0)) return LCGB200_NULL_PRECONDITION_MATRIX;
	}
	return guarded([&]() {
		HostStage<double> hs(n);
		Operator<double> A; A.h = h;
		void* user = builtin ? h->user : instance;
		auto host_call = [&](lcgb200_axfunc_ptr f, const double* x, double* y) {
			LCG_CUDA_CHECK(cudaMemcpy(hs.get(hs.x), x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
			f(user, hs.x, hs.get(hs.y), n);
			LCG_CUDA_CHECK(cudaMemcpy(y, hs.y, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
		};
		if (!builtin) { A.apply = [&](const double* x, double* y, int) { host_call(Afp, x, y); }; A.host_side = true; }
		if (solver_id == LCGB200_PCG)
		{
			if (Mfp == lcgb200_jacobi_mx_host) A.diag = (const double*)h->diag;
			else if (Mfp == lcgb200_ic0_mx_host) A.ic0 = h;
			else { A.precond = [&](const double* x, double* y, int) { host_call(Mfp, x, y); }; A.host_side = true; }
		}
		auto make_pf = [&](const double* m_dev) -> ProgressFn {
			if (!Pfp) return ProgressFn();
			return [&, m_dev](double res, int k) {
				LCG_CUDA_CHECK(cudaMemcpy(hs.get(hs.m), m_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
				return Pfp(user, hs.m, res, &para, n, k);
			};
		};
		return do_solve_real(h, A, solver_id, m, B, low, hig, para, n, h ? h->n_cols : n, h ? h->n_global : (long long)n, make_pf, false, nullptr, nullptr,
			ws_out, n_ws_out);
	});
}

}  // namespace

extern "C" {

void lcgb200_csr_ax_host(void*, const double*, double*, const int) {}
void lcgb200_jacobi_mx_host(void*, const double*, double*, const int) {}
void lcgb200_ic0_mx_host(void*, const double*, double*, const int) {}
void lcgb200_csr_cax_host(void*, const void*, void*, const int, int, int) {}

int lcgb200_solver(lcgb200_axfunc_ptr Afp, lcgb200_progress_ptr Pfp, double* m, const double* B, const int n_size,
	const lcgb200_para* param, void* instance, int solver_id)
{	// lcg.cpp:59-82: CG, CGS, BICGSTAB, BICGSTAB2; anything else -> CGS
	int id = LCGB200_CGS;
	if (solver_id == LCGB200_CG || solver_id == LCGB200_BICGSTAB || solver_id == LCGB200_BICGSTAB2) id = solver_id;
	return host_real(Afp, nullptr, Pfp, m, B, nullptr, nullptr, n_size, param, instance, id);
}

// lcg() (lcg.h:135-137, lcg.cpp:143-274): the stand-alone CG with optional caller-owned work vectors Gk, Dk, ADk.  The work
// vectors of the solve live on the device; the caller's arrays (where given) receive their final contents.
int lcgb200_lcg(lcgb200_axfunc_ptr Afp, lcgb200_progress_ptr Pfp, double* m, const double* B, const int n_size,
	const lcgb200_para* param, void* instance, double* Gk, double* Dk, double* ADk)
{
	double* ws[3] = {Gk, Dk, ADk};
	return host_real(Afp, nullptr, Pfp, m, B, nullptr, nullptr, n_size, param, instance, LCGB200_CG, ws, 3);
}

// lcgs() (lcg.h:166-169, lcg.cpp:437-612): the stand-alone CGS with optional work vectors RK, R0T, PK, AX, UK, QK, WK
int lcgb200_lcgs(lcgb200_axfunc_ptr Afp, lcgb200_progress_ptr Pfp, double* m, const double* B, const int n_size,
	const lcgb200_para* param, void* instance, double* RK, double* R0T, double* PK, double* AX, double* UK, double* QK, double* WK)
{
	double* ws[7] = {RK, R0T, PK, AX, UK, QK, WK};
	return host_real(Afp, nullptr, Pfp, m, B, nullptr, nullptr, n_size, param, instance, LCGB200_CGS, ws, 7);
}

int lcgb200_solver_preconditioned(lcgb200_axfunc_ptr Afp, lcgb200_axfunc_ptr Mfp, lcgb200_progress_ptr Pfp, double* m, const double* B,
	const int n_size, const lcgb200_para* param, void* instance, int)
{	// lcg.cpp:87-91: always lpcg
	return host_real(Afp, Mfp, Pfp, m, B, nullptr, nullptr, n_size, param, instance, LCGB200_PCG);
}

int lcgb200_solver_constrained(lcgb200_axfunc_ptr Afp, lcgb200_progress_ptr Pfp, double* m, const double* B, const double* low, const double* hig,
	const int n_size, const lcgb200_para* param, void* instance, int solver_id)
{	// lcg.cpp:121-140: PG, SPG; anything else -> PG
	return host_real(Afp, nullptr, Pfp, m, B, low, hig, n_size, param, instance, solver_id == LCGB200_SPG ? LCGB200_SPG : LCGB200_PG);
}

int lcgb200_csolver(lcgb200_caxfunc_ptr Afp, lcgb200_cprogress_ptr Pfp, void* m, const void* B, const int n_size,
	const lcgb200_cpara* param, void* instance, int solver_id)
{	// clcg.cpp:46-74: BICG, BICG_SYM, CGS, BICGSTAB, TFQMR; anything else -> CGS
	const lcgb200_cpara para = param ? *param : kDefCPara;
	const int n = n_size;
	int rc = check_cplx(n, para, m, B);
	if (rc) return rc;
	if (!Afp) return LCGB200_C_INVALID_POINTER;
	if (solver_id < LCGB200_CBICG || solver_id > LCGB200_CTFQMR) solver_id = LCGB200_CCGS;
	const bool builtin = (Afp == lcgb200_csr_cax_host);
	CsrHandle* h = builtin ? reinterpret_cast<CsrHandle*>(instance) : nullptr;
	if (builtin)
	{
		if (!h) return LCGB200_C_INVALID_POINTER;
		if (h->value_type != LCGB200_COMPLEX || h->n_rows != n) return LCGB200_C_SIZE_NOT_MATCH;
		if (solver_id == LCGB200_CBICG && !h->t_row_ptr && !h->t_handle) { set_error_msg("CLCG_BICG needs LCGB200_CSR_TRANSPOSE"); return LCGB200_C_UNKNOWN_SOLVER; }
	}
	return guarded([&]() {
		HostStage<double2> hs(n);
		Operator<double2> A; A.h = h;
		void* user = builtin ? h->user : instance;
		A.host_side = !builtin;
		if (!builtin)
			A.apply = [&](const double2* x, double2* y, int op) {	// op 0 = A x, 1 = A^T x, 2 = A^H x  ->  (layout, conjugate) of clcg.h:40-41
				LCG_CUDA_CHECK(cudaMemcpy(hs.get(hs.x), x, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost));
				Afp(user, hs.x, hs.get(hs.y), n, op == 0 ? 0 : 1, op == 2 ? 1 : 0);
				LCG_CUDA_CHECK(cudaMemcpy(y, hs.y, sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice));
			};
		auto make_pf = [&](const double2* m_dev) -> ProgressFn {
			if (!Pfp) return ProgressFn();
			return [&, m_dev](double res, int k) {
				LCG_CUDA_CHECK(cudaMemcpy(hs.get(hs.m), m_dev, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost));
				return Pfp(user, hs.m, res, &para, n, k);
			};
		};
		return do_solve_cplx<double2>(h, A, solver_id, (double2*)m, (const double2*)B, para, n, h ? std::max(h->n_cols, h->t_handle ? h->t_handle->n_cols : 0) : n, h ? h->n_global : (long long)n, make_pf, false, nullptr, nullptr);
	});
}

}  // extern "C"

