// CPU check of the row-pattern operator copy's host logic (liblcg_b200/csrc/pat_host.h) and of the tiling k_spmv_pat uses
// (liblcg_b200/csrc/csr.cuh "row-pattern operator"): for a set of stencil matrices — cubes, bricks whose nx is not a multiple
// of 32, a row block with ghost columns, a matrix without any repeating offset — it
//   1. finds the distinct rows (patterns), picks the stride S and builds the chains exactly as capi.cu: try_patterns does,
//   2. walks the work items of both kernels with their index arithmetic — k_spmv_pat: item flag, chained path with
//      R + m - 1 loads per chain, masked path (rows whose patterns are subsets of one longer pattern), row-by-row path;
//      k_spmv_pat_march (when the plan exists): segments, block items, the producer's clipped window lines (what it does
//      not copy stays stale = NaN here), the consumers' reads through (plane, line shift, position); k_spmv_pat_box (dense box
//      stencils): the flag byte of every thread and the pairs / side values it reads under those flags — and
//   3. checks that every row is written exactly once, that no read leaves [0, n_cols) or returns a stale value, that the
//      bulk copies are 16-byte aligned, and that y equals the CSR product to rounding.
// No GPU, no CUDA: g++ -O2 -std=c++17 -I liblcg_b200/csrc tests/cxx/pat_chain_check.cpp.  Prints one line per case.
#include "pat_host.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <string>

using namespace lcgb200;
constexpr int R = 8;   // csr.cuh: kPatRows

struct Csr { int n_rows = 0, n_cols = 0; std::vector<int> rp, ci; std::vector<double> v; };

// 7- or 27-point stencil on an nx x ny x nz brick; rows [z0, z1) planes only (a row block), columns outside the block
// remapped behind the local ones ("ghosts"), as liblcg_b200/dist.py does for a z-slab partition
static Csr stencil(int nx, int ny, int nz, bool full27, int z0, int z1, int coef_period = 1)
{
	Csr A;
	const long long plane = (long long)nx * ny;
	const int n_loc = (int)(plane * (z1 - z0));
	const int lo_ghost = z0 > 0 ? (int)plane : 0, hi_ghost = z1 < nz ? (int)plane : 0;
	A.n_rows = n_loc; A.n_cols = n_loc + lo_ghost + hi_ghost;
	A.rp.push_back(0);
	for (int k = z0; k < z1; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++)
	{
		for (int dk = -1; dk <= 1; dk++) for (int dj = -1; dj <= 1; dj++) for (int di = -1; di <= 1; di++)
		{
			const int nzero = (di != 0) + (dj != 0) + (dk != 0);
			if (!full27 && nzero > 1) continue;
			const int ii = i + di, jj = j + dj, kk = k + dk;
			if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) continue;
			long long col;
			if (kk < z0) col = n_loc + (long long)jj * nx + ii;
			else if (kk >= z1) col = n_loc + lo_ghost + (long long)jj * nx + ii;
			else col = ((long long)(kk - z0) * ny + jj) * nx + ii;
			A.ci.push_back((int)col);
			// coef_period > 1: the coefficients change from grid line to grid line, so rows S apart hold different values and
			// are no subsets of one another (the row-by-row path)
			A.v.push_back((nzero == 0 ? (full27 ? 26.0 : 6.0) + 0.5 : -1.0 / (1 + nzero)) * (1.0 + 0.25 * (j % coef_period)));
		}
		A.rp.push_back((int)A.ci.size());
	}
	return A;
}

// offsets {-1, 0, +1}: no difference >= 32, the stride falls back to the default and every chain has one entry
static Csr tridiagonal(int n)
{
	Csr A; A.n_rows = A.n_cols = n; A.rp.push_back(0);
	for (int i = 0; i < n; i++)
	{
		if (i > 0) { A.ci.push_back(i - 1); A.v.push_back(-1.0); }
		A.ci.push_back(i); A.v.push_back(2.5);
		if (i + 1 < n) { A.ci.push_back(i + 1); A.v.push_back(-1.25); }
		A.rp.push_back((int)A.ci.size());
	}
	return A;
}

struct Setup {
	int S = 0; size_t longest = 0;
	std::vector<std::vector<std::pair<int, double>>> rows;
	std::vector<int> pat, info, sup;
	std::vector<unsigned long long> mask;
	std::vector<std::vector<PatChainH>> chains;
	PatMarchH plan;
	PatBoxH box; std::vector<unsigned char> rowflags;
};

// capi.cu: try_patterns, host part
static bool setup(const std::string& name, const Csr& A, Setup& T)
{
	const int n = A.n_rows;
	std::map<std::vector<std::pair<int, double>>, int> ids;
	T.pat.assign((size_t)n, 0);
	for (int r = 0; r < n; r++)
	{
		std::vector<std::pair<int, double>> e;
		for (int k = A.rp[(size_t)r]; k < A.rp[(size_t)r + 1]; k++) e.push_back({A.ci[(size_t)k] - r, A.v[(size_t)k]});
		auto f = ids.find(e);
		if (f == ids.end()) { f = ids.insert({e, (int)T.rows.size()}).first; T.rows.push_back(e); }
		T.pat[(size_t)r] = f->second;
	}
	const size_t np = T.rows.size();
	if (np > 253) { std::printf("%s: %zu patterns (> 253)\n", name.c_str(), np); return false; }
	// the geometry pattern: the longest row, the most frequent one among equally long rows (capi.cu: k_pat_count)
	std::vector<long long> count(np, 0);
	for (int r = 0; r < n; r++) count[(size_t)T.pat[(size_t)r]]++;
	for (size_t p = 0; p < np; p++)
		if (T.rows[p].size() > T.rows[T.longest].size() || (T.rows[p].size() == T.rows[T.longest].size() && count[p] > count[T.longest])) T.longest = p;
	T.S = pat_pick_stride(T.rows[T.longest], n, R, 256);
	T.chains.resize(np); T.info.resize(np);
	std::vector<int> t0s(np, -1);
	for (size_t p = 0; p < np; p++)
	{
		pat_build_chains(T.rows[p], T.S, T.chains[p], &t0s[p]);
		T.info[p] = (int)T.chains[p].size() | ((t0s[p] + 1) << 8);
		size_t entries = 0;
		for (const auto& c : T.chains[p]) entries += (size_t)c.m;
		if (entries != T.rows[p].size()) { std::printf("%s: pattern %zu lost entries\n", name.c_str(), p); return false; }
		if (t0s[p] >= 0 && (long long)T.chains[p].back().off + (long long)t0s[p] * T.S != 0) { std::printf("%s: the diagonal's chain is not the last one\n", name.c_str()); return false; }
	}
	pat_plan_march(T.chains[T.longest], t0s[T.longest], T.S, n, A.n_cols, R, T.plan);
	if (T.plan.ok)
	{	// the plan must keep the diagonal chain last and describe every chain's offset through (plane, shift, position)
		const auto& ch = T.chains[T.longest];
		if (t0s[T.longest] >= 0 && (long long)ch.back().off + (long long)t0s[T.longest] * T.S != 0) { std::printf("%s: plan moved the diagonal chain\n", name.c_str()); return false; }
		for (const auto& c : ch)
		{
			const int sh = (c.m >> 4) & 15, d = (c.m >> 8) & 255, pl = c.m >> 16;
			if ((long long)T.plan.o0 + (long long)pl * T.plan.S2 + (long long)sh * T.S + d != c.off || d > kPatSpanH || sh > kPatMaxShiftH) { std::printf("%s: bad placement\n", name.c_str()); return false; }
			if (T.plan.wy * R + sh + (c.m & 3) - 1 > T.plan.nlines) { std::printf("%s: window too short\n", name.c_str()); return false; }
		}
	}
	pat_build_masks(T.rows, T.chains, T.S, T.sup, T.mask);
	if (!T.plan.ok && T.S % 2 == 0) pat_plan_box(T.chains[T.longest], T.S, (int)T.longest, T.sup, T.mask, T.box, T.rowflags);
	return true;
}

struct Result { std::vector<double> y, xc; std::vector<int> written; long long reads = 0; int bad = 0; };

// csr.cuh: pat_item_ldg for one lane (one thread's R rows)
static void item_ldg(const Csr& A, const Setup& T, const std::vector<double>& x, long long row0, int uni, Result& res)
{
	const int n = A.n_rows, S = T.S;
	auto ld = [&](long long idx) -> double { res.reads++; if (idx < 0 || idx >= A.n_cols) { res.bad++; return 0.0; } return x[(size_t)idx]; };
	int p = uni; bool chained = true, masked = false;
	unsigned long long mk[R];
	if (uni == 255)
	{
		int p0 = -1, sp = -1;
		masked = true;
		for (int q = 0; q < R; q++)
		{
			const long long row = row0 + (long long)q * S;
			const int pq = row < n ? T.pat[(size_t)row] : -1;
			if (q == 0) p0 = pq;
			chained = chained && pq >= 0 && pq == p0;
			mk[q] = 0ull;
			if (pq >= 0)
			{
				if (sp < 0) sp = T.sup[(size_t)pq];
				masked = masked && T.sup[(size_t)pq] == sp && T.mask[(size_t)pq] != 0ull;
				mk[q] = T.mask[(size_t)pq];
			}
		}
		p = chained ? p0 : sp;
	}
	auto store = [&](long long row, double v, double xc) { res.y[(size_t)row] = v; res.written[(size_t)row]++; res.xc[(size_t)row] = xc; };
	if (chained)
	{
		const int nch = T.info[(size_t)p] & 255, t0 = (T.info[(size_t)p] >> 8) - 1;
		double sum[R];
		for (int q = 0; q < R; q++) sum[q] = 0.0;
		for (int c = 0; c < nch; c++)
		{
			const PatChainH& ch = T.chains[(size_t)p][(size_t)c];
			const int m = ch.m & 3;
			const long long xb = row0 + ch.off;
			double xl[R + 2];
			for (int u = 0; u < R; u++) xl[u] = ld(xb + (long long)u * S);
			xl[R] = m > 1 ? ld(xb + (long long)R * S) : 0.0;
			xl[R + 1] = m > 2 ? ld(xb + (long long)(R + 1) * S) : 0.0;
			for (int t = 0; t < m; t++) for (int q = 0; q < R; q++) sum[q] = std::fma(ch.v[t], xl[q + t], sum[q]);
			if (c == nch - 1 && t0 >= 0) for (int q = 0; q < R; q++) store(row0 + (long long)q * S, sum[q], xl[q + t0]);
		}
		if (t0 < 0) for (int q = 0; q < R; q++) store(row0 + (long long)q * S, sum[q], ld(row0 + (long long)q * S));
	}
	else if (masked && p >= 0)
	{
		const int nch = T.info[(size_t)p] & 255;
		double sum[R];
		for (int q = 0; q < R; q++) sum[q] = 0.0;
		for (int c = 0; c < nch; c++)
		{
			const PatChainH& ch = T.chains[(size_t)p][(size_t)c];
			const long long xb = row0 + ch.off;
			unsigned int b[R + 2];
			for (int q = 0; q < R; q++) b[q] = (unsigned int)(mk[q] >> (3 * c)) & 7u;
			b[R] = 0u; b[R + 1] = 0u;
			double xl[R + 2];
			for (int u = 0; u < R + 2; u++)
			{
				const unsigned int need = (b[u] & 1u) | (u >= 1 ? (b[u - 1] & 2u) : 0u) | (u >= 2 ? (b[u - 2] & 4u) : 0u);
				xl[u] = need ? ld(xb + (long long)u * S) : 0.0;
			}
			for (int q = 0; q < R; q++) for (int t = 0; t < 3; t++) if (b[q] & (1u << t)) sum[q] = std::fma(ch.v[t], xl[q + t], sum[q]);
		}
		for (int q = 0; q < R; q++) { const long long row = row0 + (long long)q * S; if (row < n) store(row, sum[q], ld(row)); }
	}
	else
	{
		for (int q = 0; q < R; q++)
		{
			const long long row = row0 + (long long)q * S;
			if (row >= n) break;
			const int pq = T.pat[(size_t)row];
			double sum = 0.0;
			for (const PatChainH& ch : T.chains[(size_t)pq]) for (int t = 0; t < (ch.m & 3); t++) sum = std::fma(ch.v[t], ld(row + ch.off + (long long)t * S), sum);
			store(row, sum, ld(row));
		}
	}
}

// csr.cuh: k_spmv_pat (+ capi.cu: k_pat_items)
static void walk_ldg(const Csr& A, const Setup& T, const std::vector<double>& x, Result& res)
{
	const int n = A.n_rows, S = T.S;
	const long long n_super = ((long long)n + S - 1) / S, n_a = (n_super + R - 1) / R;
	const int nib = (S + 31) / 32, n_items = (int)(n_a * nib);
	for (int it = 0; it < n_items; it++)
	{
		const int a = it / nib, ib = it - a * nib;
		int first = -2; bool any_bad = false;
		for (int lane = 0; lane < 32; lane++)
		{
			const int i = ib * 32 + lane;
			if (i >= S) continue;
			for (int q = 0; q < R; q++)
			{
				const long long row = ((long long)a * R + q) * S + i;
				const int pq = row < n ? T.pat[(size_t)row] : -1;
				if (first == -2) first = pq;
				any_bad = any_bad || pq < 0 || pq != first;
			}
		}
		const int uni = any_bad ? 255 : first;
		for (int lane = 0; lane < 32; lane++)
		{
			const int i = ib * 32 + lane;
			if (i < S) item_ldg(A, T, x, (long long)a * R * S + i, uni, res);
		}
	}
}

// csr.cuh: k_spmv_pat_march (+ capi.cu: k_pat_bitems, pat_host.h: pat_build_segments).  The mbarrier ring itself is not
// emulated; what is: which window every read goes to, what the producer put there (clipped lines, the rest stale = NaN),
// and that every window index a consumer touches lies in [k, k + G).
static void walk_march(const Csr& A, const Setup& T, const std::vector<double>& x, int target, Result& res, long long* fallback_items)
{
	const int n = A.n_rows, S = T.S;
	const PatMarchH& M = T.plan;
	const int WX = M.wx, WY = M.wy, WD = 32 * WX + kPatSpanH;
	const int gpat = (int)T.longest;
	const int t0 = (T.info[(size_t)gpat] >> 8) - 1, c_last = (T.info[(size_t)gpat] & 255) - 1;
	std::vector<PatSegH> segs;
	pat_build_segments(M, n, S, R, target, 2, segs);
	const double STALE = std::nan("");
	auto store = [&](long long row, double v, double xc) { res.y[(size_t)row] = v; res.written[(size_t)row]++; res.xc[(size_t)row] = xc; };
	for (const PatSegH& seg : segs)
	{
		const long long wbase0 = (long long)seg.ab * WY * R * S + (long long)seg.ibb * (32 * WX) + M.o0;
		auto window = [&](int j) {   // producer
			std::vector<double> w((size_t)M.nlines * WD, STALE);
			const long long wbase = wbase0 + (long long)j * M.S2;
			for (int l = 0; l < M.nlines; l++)
			{
				const long long b0 = wbase + (long long)l * S;
				const long long lo = b0 < 0 ? 0 : b0, hi = b0 + WD > A.n_cols ? A.n_cols : b0 + WD;
				if ((lo & 1) || (hi & 1)) res.bad++;   // 16-byte alignment of the bulk copies
				for (long long e = lo; e < hi; e++) { w[(size_t)(l * WD + (e - b0))] = x[(size_t)e]; res.reads++; }
			}
			return w;
		};
		for (int k = 0; k < seg.len; k++)
		{
			std::vector<std::vector<double>> wins;
			for (int p = 0; p < M.G; p++) wins.push_back(window(k + p));
			const long long ab = seg.ab + (long long)k * M.dAb;
			// k_pat_bitems
			int uni_w[8]; bool geo_bad = false;
			for (int w = 0; w < 8; w++)
			{
				const int wq = w / WX, wi = w % WX;
				int first = -2; bool mixed = false, any_row = false;
				for (int lane = 0; lane < 32; lane++)
					for (int q = 0; q < R; q++)
					{
						const long long row = ((ab * WY + wq) * R + q) * S + (long long)(seg.ibb * WX + wi) * 32 + lane;
						const int pq = row < n ? T.pat[(size_t)row] : -1;
						if (first == -2) first = pq;
						mixed = mixed || pq != first || pq < 0;
						if (pq >= 0) { any_row = true; geo_bad = geo_bad || T.sup[(size_t)pq] != gpat || T.mask[(size_t)pq] == 0ull; }
					}
				uni_w[w] = !any_row ? 254 : (mixed ? 255 : first);
			}
			if (geo_bad) (*fallback_items)++;
			for (int w = 0; w < 8; w++)
			{
				const int wq = w / WX, wi = w % WX, uni = uni_w[w];
				if (uni == 254) continue;
				for (int lane = 0; lane < 32; lane++)
				{
					const long long row0 = ((ab * WY + wq) * R) * S + (long long)(seg.ibb * WX + wi) * 32 + lane;
					if (geo_bad) { item_ldg(A, T, x, row0, uni, res); continue; }
					unsigned long long mq[R]; bool same = true;
					for (int q = 0; q < R; q++)
					{
						const long long row = row0 + (long long)q * S;
						mq[q] = uni != 255 ? T.mask[(size_t)uni] : (row < n ? T.mask[(size_t)T.pat[(size_t)row]] : 0ull);
						same = same && mq[q] == mq[0];
					}
					double sum[R];
					for (int q = 0; q < R; q++) sum[q] = 0.0;
					for (int g = 0; g < M.G; g++)
					{
						const std::vector<double>& win = wins[(size_t)M.group_plane[g]];
						for (int c = M.group_begin[g]; c < M.group_begin[g + 1]; c++)
						{
							const PatChainH& ch = T.chains[(size_t)gpat][(size_t)c];
							const int m = ch.m & 3, sh = (ch.m >> 4) & 15, d = (ch.m >> 8) & 255;
							if ((ch.m >> 16) != M.group_plane[g]) res.bad++;
							double xl[R + 2];
							for (int u = 0; u < R + 2; u++)
							{
								const int line = wq * R + sh + u;
								xl[u] = (u < R || m > u - R + 1) ? win[(size_t)(line * WD + d + wi * 32 + lane)] : 0.0;
							}
							for (int q = 0; q < R; q++)
							{
								const unsigned int pres = (unsigned int)((same ? mq[0] : mq[q]) >> (3 * c)) & 7u;
								for (int t = 0; t < 3; t++)
									if (pres & (1u << t))
									{
										if (std::isnan(xl[q + t])) res.bad++;   // a value the producer never copied
										sum[q] = std::fma(ch.v[t], xl[q + t], sum[q]);
									}
							}
							if (c == c_last && t0 >= 0)
								for (int q = 0; q < R; q++)
								{
									const long long row = row0 + (long long)q * S;
									if (uni != 255 || row < n) store(row, sum[q], xl[q + t0]);
								}
						}
					}
					if (t0 < 0)
						for (int q = 0; q < R; q++)
						{
							const long long row = row0 + (long long)q * S;
							if (uni != 255 || row < n) store(row, sum[q], x[(size_t)row]);
						}
				}
			}
		}
	}
}

// csr.cuh: k_spmv_pat_box (+ capi.cu: k_pat_box_flags).  *odd = warps that hold a thread (with rows) that falls back to the chain tables
static void walk_box(const Csr& A, const Setup& T, const std::vector<double>& x, Result& res, long long* odd)
{
	const int n = A.n_rows, S = T.S;
	const PatBoxH& B = T.box;
	const long long n_super = ((long long)n + S - 1) / S, n_a = (n_super + R - 1) / R;
	const int nib = (S + 63) / 64;
	auto ld = [&](long long idx) -> double { res.reads++; if (idx < 0 || idx >= A.n_cols) { res.bad++; return 0.0; } return x[(size_t)idx]; };
	auto store = [&](long long row, double v, double xc) { res.y[(size_t)row] = v; res.written[(size_t)row]++; res.xc[(size_t)row] = xc; };
	for (long long a = 0; a < n_a; a++)
		for (int ib = 0; ib < nib; ib++)
		{
			bool warp_odd = false;
			for (int lane = 0; lane < 32; lane++)
			{
				const int i = ib * 64 + 2 * lane;
				if (i >= S) continue;
				const long long row0 = a * R * S + i;
				// k_pat_box_flags
				int out = 0, planes = -1; bool bad = false, any_row = false;
				for (int q = 0; q < R; q++)
					for (int e = 0; e < 2; e++)
					{
						const long long row = row0 + (long long)q * S + e;
						if (row >= n) { bad = true; continue; }
						any_row = true;
						const int f = T.rowflags[(size_t)T.pat[(size_t)row]];
						if (f & 0x80) { bad = true; continue; }
						if ((f & kBoxHDropL) && e != 0) bad = true;
						if ((f & kBoxHDropR) && e != 1) bad = true;
						if ((f & kBoxHDropLow) && q != 0) bad = true;
						if ((f & kBoxHDropHigh) && q != R - 1) bad = true;
						if (planes < 0) planes = f & ~15;
						if ((f & ~15) != planes) bad = true;
						out |= f;
					}
				if (!bad)
					for (int q = 0; q < R && !bad; q++)
						for (int e = 0; e < 2; e++)
						{
							const int f = T.rowflags[(size_t)T.pat[(size_t)(row0 + (long long)q * S + e)]];
							if ((out & kBoxHDropL) && e == 0 && !(f & kBoxHDropL)) bad = true;
							if ((out & kBoxHDropR) && e == 1 && !(f & kBoxHDropR)) bad = true;
							if ((out & kBoxHDropLow) && q == 0 && !(f & kBoxHDropLow)) bad = true;
							if ((out & kBoxHDropHigh) && q == R - 1 && !(f & kBoxHDropHigh)) bad = true;
						}
				const int flags = bad ? 255 : out;
				if (bad && any_row) warp_odd = true;
				// k_spmv_pat_box
				if (flags != 255)
				{
					double s0[R], s1[R];
					for (int q = 0; q < R; q++) { s0[q] = 0.0; s1[q] = 0.0; }
					for (int g = 0; g < B.G; g++)
					{
						if (flags & (kBoxHDropG0 << g)) continue;
						const long long xg = row0 + B.center[g];
						if (xg & 1) res.bad++;   // 16-byte alignment of the pair loads
						for (int u = 0; u < R + 2; u++)
						{
							double xm = 0.0, x0 = 0.0, x1 = 0.0, xp = 0.0;
							const bool line_on = !((u == 0 && (flags & kBoxHDropLow)) || (u == R + 1 && (flags & kBoxHDropHigh)));
							if (line_on)
							{
								const long long pl = xg + (long long)u * S;
								x0 = ld(pl); x1 = ld(pl + 1);
								if (!(flags & kBoxHDropL)) xm = ld(pl - 1);
								if (!(flags & kBoxHDropR)) xp = ld(pl + 2);
							}
							for (int j = 0; j < 3; j++)
							{
								const int q = u - j;
								if (q < 0 || q >= R) continue;
								s0[q] = std::fma(B.coef[g][0][j], xm, s0[q]); s0[q] = std::fma(B.coef[g][1][j], x0, s0[q]); s0[q] = std::fma(B.coef[g][2][j], x1, s0[q]);
								s1[q] = std::fma(B.coef[g][0][j], x0, s1[q]); s1[q] = std::fma(B.coef[g][1][j], x1, s1[q]); s1[q] = std::fma(B.coef[g][2][j], xp, s1[q]);
							}
						}
					}
					for (int q = 0; q < R; q++)
					{
						const long long row = row0 + (long long)q * S;
						store(row, s0[q], ld(row)); store(row + 1, s1[q], ld(row + 1));
					}
				}
				else
					for (int e = 0; e < 2; e++) item_ldg(A, T, x, row0 + e, 255, res);
			}
			if (warp_odd) (*odd)++;
		}
}

static int verify(const std::string& name, const char* which, const Csr& A, const std::vector<double>& x, const Result& res)
{
	const int n = A.n_rows;
	double err = 0.0, scale = 0.0;
	int once = 0, xc_bad = 0;
	for (int r = 0; r < n; r++)
	{
		double s = 0.0;
		for (int k = A.rp[(size_t)r]; k < A.rp[(size_t)r + 1]; k++) s = std::fma(A.v[(size_t)k], x[(size_t)A.ci[(size_t)k]], s);
		err = std::max(err, std::fabs(s - res.y[(size_t)r])); scale = std::max(scale, std::fabs(s));
		once += res.written[(size_t)r] == 1;
		xc_bad += !(res.xc[(size_t)r] == x[(size_t)r]);
	}
	std::printf("    %-6s reads/nnz=%.3f err=%.2e\n", which, (double)res.reads / (double)A.rp[(size_t)n], err / scale);
	int fail = 0;
	if (once != n) { std::printf("  %s %s: rows written exactly once: %d of %d\n", name.c_str(), which, once, n); fail = 1; }
	if (res.bad) { std::printf("  %s %s: %d reads out of range / of values never staged / misaligned copies\n", name.c_str(), which, res.bad); fail = 1; }
	if (xc_bad) { std::printf("  %s %s: %d rows handed a wrong x[row] to the epilogue\n", name.c_str(), which, xc_bad); fail = 1; }
	if (!(err / scale < 1e-14)) { std::printf("  %s %s: y differs from the CSR product\n", name.c_str(), which); fail = 1; }
	return fail;
}

// expect_march: 1 = the plan must exist, 0 = must not, -1 = either
static int check(const std::string& name, const Csr& A, int expect_stride, size_t expect_chains, int expect_march = -1)
{
	Setup T;
	if (!setup(name, A, T)) return 1;
	const int n = A.n_rows;
	std::mt19937_64 rng(7);
	std::uniform_real_distribution<double> U(-1.0, 1.0);
	std::vector<double> x((size_t)A.n_cols);
	for (auto& t : x) t = U(rng);
	std::printf("%-30s n=%-8d patterns=%-3zu S=%-5d chains=%-2zu", name.c_str(), n, T.rows.size(), T.S, T.chains[T.longest].size());
	if (T.plan.ok) std::printf(" march: G=%d S2=%d o0=%d lines=%d wx=%d dAb=%d", T.plan.G, T.plan.S2, T.plan.o0, T.plan.nlines, T.plan.wx, T.plan.dAb);
	std::printf("\n");
	int fail = 0;
	Result r1; r1.y.assign((size_t)n, 0.0); r1.xc.assign((size_t)n, 0.0); r1.written.assign((size_t)n, 0);
	walk_ldg(A, T, x, r1);
	fail |= verify(name, "ldg", A, x, r1);
	if (T.plan.ok)
		for (int target : {1, 7})
		{
			Result r2; r2.y.assign((size_t)n, 0.0); r2.xc.assign((size_t)n, 0.0); r2.written.assign((size_t)n, 0);
			long long fb = 0;
			walk_march(A, T, x, target, r2, &fb);
			fail |= verify(name, target == 7 ? "march7" : "march", A, x, r2);
			if (target == 7 && fb) std::printf("    %lld block items through the plain-load path\n", fb);
		}
	if (T.box.ok)
	{
		Result r3; r3.y.assign((size_t)n, 0.0); r3.xc.assign((size_t)n, 0.0); r3.written.assign((size_t)n, 0);
		long long odd = 0;
		walk_box(A, T, x, r3, &odd);
		fail |= verify(name, "box", A, x, r3);
		const double n_warps = (double)(((n + T.S - 1) / T.S + R - 1) / R) * ((T.S + 63) / 64);
		std::printf("    box: %d planes, %.1f %% of the warps hold a thread that falls back to the chain tables%s\n", T.box.G, 100.0 * (double)odd / n_warps,
			(double)odd <= 0.10 * n_warps ? "" : " (kernel not used)");
	}
	if (expect_stride && T.S != expect_stride) { std::printf("  expected stride %d\n", expect_stride); fail = 1; }
	if (expect_chains && T.chains[T.longest].size() != expect_chains) { std::printf("  expected %zu chains\n", expect_chains); fail = 1; }
	if (expect_march >= 0 && T.plan.ok != expect_march) { std::printf("  march plan: expected %d\n", expect_march); fail = 1; }
	return fail;
}

int main()
{
	int fail = 0;
	fail |= check("27pt 256x8x6", stencil(256, 8, 6, true, 0, 6), 256, 9, 1);
	fail |= check("27pt 128x16x7", stencil(128, 16, 7, true, 0, 7), 128, 9, 1);
	fail |= check("27pt 128x32x5 (dAb = 2)", stencil(128, 32, 5, true, 0, 5), 128, 9, 1);
	fail |= check("27pt 256x24x4 (dAb = 3)", stencil(256, 24, 4, true, 0, 4), 256, 9, 1);
	fail |= check("7pt 128x16x9", stencil(128, 16, 9, false, 0, 9), 128, 5, 1);
	fail |= check("7pt 256x16x5", stencil(256, 16, 5, false, 0, 5), 256, 5, 1);
	fail |= check("27pt 128x16x12 slab z=[3,8)", stencil(128, 16, 12, true, 3, 8), 128, 9, 1);
	fail |= check("7pt 256x8x9 slab z=[0,4)", stencil(256, 8, 9, false, 0, 4), 256, 5, 1);
	fail |= check("27pt 128x12x6 (ny % 8 != 0)", stencil(128, 12, 6, true, 0, 6), 128, 9, 0);
	fail |= check("27pt 64^3", stencil(64, 64, 64, true, 0, 64), 64, 9, 0);
	fail |= check("27pt 44^3", stencil(44, 44, 44, true, 0, 44), 44, 9);
	fail |= check("27pt 64x16x40 slab z=[6,34)", stencil(64, 16, 40, true, 6, 34), 64, 9, 0);   // box kernel with two ghost-coupled planes
	fail |= check("27pt 50x37x29", stencil(50, 37, 29, true, 0, 29), 50, 9);
	fail |= check("7pt 50^3", stencil(50, 50, 50, false, 0, 50), 50, 5);
	fail |= check("7pt 96x40x33", stencil(96, 40, 33, false, 0, 33), 96, 5);
	fail |= check("27pt 48^3 slab z=[12,24)", stencil(48, 48, 48, true, 12, 24), 48, 9);
	fail |= check("7pt 40^3 slab z=[0,13)", stencil(40, 40, 40, false, 0, 13), 40, 5);
	fail |= check("27pt 33x5x4 (tiny)", stencil(33, 5, 4, true, 0, 4), 0, 0);
	fail |= check("27pt 8^3 (nx < 32)", stencil(8, 8, 8, true, 0, 8), 64, 9);
	fail |= check("7pt 40^3 coefficients j%3", stencil(40, 40, 40, false, 0, 40, 3), 40, 5);
	fail |= check("tridiagonal 5000", tridiagonal(5000), 256, 3);
	fail |= check("tridiagonal 100", tridiagonal(100), 256, 3);
	std::printf(fail ? "FAILED\n" : "OK\n");
	return fail;
}
