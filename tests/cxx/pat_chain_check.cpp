// CPU check of the row-pattern operator copy's host logic (liblcg_b200/csrc/pat_host.h) and of the tiling k_spmv_pat uses
// (liblcg_b200/csrc/csr.cuh "row-pattern operator"): for a set of stencil matrices — cubes, bricks whose nx is not a multiple
// of 32, a row block with ghost columns, a matrix without any repeating offset — it
//   1. finds the distinct rows (patterns), picks the stride S and builds the chains exactly as capi.cu: try_patterns does,
//   2. walks the work items (a, ib, lane, q) with the kernel's index arithmetic: item flag, chained path with R + m - 1
//      loads per chain, masked path (rows whose patterns are subsets of one longer pattern), row-by-row path, and
//   3. checks that every row is written exactly once, that no load leaves [0, n_cols), and that y equals the CSR product
//      to rounding.
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

static int check(const std::string& name, const Csr& A, int expect_stride, size_t expect_chains)
{
	const int n = A.n_rows;
	// 1. patterns
	std::map<std::vector<std::pair<int, double>>, int> ids;
	std::vector<std::vector<std::pair<int, double>>> rows;
	std::vector<int> pat((size_t)n);
	for (int r = 0; r < n; r++)
	{
		std::vector<std::pair<int, double>> e;
		for (int k = A.rp[(size_t)r]; k < A.rp[(size_t)r + 1]; k++) e.push_back({A.ci[(size_t)k] - r, A.v[(size_t)k]});
		auto f = ids.find(e);
		if (f == ids.end()) { f = ids.insert({e, (int)rows.size()}).first; rows.push_back(e); }
		pat[(size_t)r] = f->second;
	}
	if (rows.size() > 255) { std::printf("%s: %zu patterns (> 255)\n", name.c_str(), rows.size()); return 1; }
	size_t longest = 0;
	for (size_t p = 0; p < rows.size(); p++) if (rows[p].size() > rows[longest].size()) longest = p;
	const int S = pat_pick_stride(rows[longest], n, R, 256);
	std::vector<std::vector<PatChainH>> chains(rows.size());
	std::vector<int> info(rows.size());
	for (size_t p = 0; p < rows.size(); p++)
	{
		int t0 = -1;
		pat_build_chains(rows[p], S, chains[p], &t0);
		info[p] = (int)chains[p].size() | ((t0 + 1) << 8);
		size_t entries = 0;
		for (const auto& c : chains[p]) entries += (size_t)c.m;
		if (entries != rows[p].size()) { std::printf("%s: pattern %zu lost entries\n", name.c_str(), p); return 1; }
		if (t0 >= 0 && (long long)chains[p][0].off + (long long)t0 * S != 0) { std::printf("%s: t0 wrong\n", name.c_str()); return 1; }
	}
	std::vector<int> sup; std::vector<unsigned long long> mask;
	pat_build_masks(rows, chains, S, sup, mask);
	// 2. the kernel's walk
	const long long n_super = ((long long)n + S - 1) / S, n_a = (n_super + R - 1) / R;
	const int nib = (S + 31) / 32, n_items = (int)(n_a * nib);
	std::mt19937_64 rng(7);
	std::uniform_real_distribution<double> U(-1.0, 1.0);
	std::vector<double> x((size_t)A.n_cols), y((size_t)n, 0.0), xc_seen((size_t)n, 0.0);
	for (auto& t : x) t = U(rng);
	std::vector<int> written((size_t)n, 0);
	long long loads = 0, fast_items = 0, chained_threads = 0, slow_rows = 0, masked_rows = 0;
	int bad = 0;
	auto ld = [&](long long idx) -> double { loads++; if (idx < 0 || idx >= A.n_cols) { bad++; return 0.0; } return x[(size_t)idx]; };
	for (int it = 0; it < n_items; it++)
	{
		const int a = it / nib, ib = it - a * nib;
		// capi.cu: k_pat_items
		int first = -2; bool any_bad = false;
		for (int lane = 0; lane < 32; lane++)
		{
			const int i = ib * 32 + lane;
			if (i >= S) continue;
			for (int q = 0; q < R; q++)
			{
				const long long row = ((long long)a * R + q) * S + i;
				const int pq = row < n ? pat[(size_t)row] : -1;
				if (first == -2) first = pq;
				any_bad = any_bad || pq < 0 || pq != first;
			}
		}
		const int uni = any_bad ? 255 : first;
		if (uni != 255) fast_items++;
		// csr.cuh: k_spmv_pat
		for (int lane = 0; lane < 32; lane++)
		{
			const int i = ib * 32 + lane;
			if (i >= S) continue;
			const long long row0 = (long long)a * R * S + i;
			int p = uni; bool chained = true, masked = false;
			unsigned long long mk[R];
			if (uni == 255)
			{
				int p0 = -1, sp = -1;
				masked = true;
				for (int q = 0; q < R; q++)
				{
					const long long row = row0 + (long long)q * S;
					const int pq = row < n ? pat[(size_t)row] : -1;
					if (q == 0) p0 = pq;
					chained = chained && pq >= 0 && pq == p0;
					mk[q] = 0ull;
					if (pq >= 0)
					{
						if (sp < 0) sp = sup[(size_t)pq];
						masked = masked && sup[(size_t)pq] == sp && mask[(size_t)pq] != 0ull;
						mk[q] = mask[(size_t)pq];
					}
				}
				p = chained ? p0 : sp;
			}
			if (chained)
			{
				chained_threads++;
				const int nch = info[(size_t)p] & 255, t0 = (info[(size_t)p] >> 8) - 1;
				double sum[R], xc[R];
				for (int q = 0; q < R; q++) { sum[q] = 0.0; xc[q] = 0.0; }
				for (int c = 0; c < nch; c++)
				{
					const PatChainH& ch = chains[(size_t)p][(size_t)c];
					const long long xb = row0 + ch.off;
					double xl[R + 2];
					for (int u = 0; u < R; u++) xl[u] = ld(xb + (long long)u * S);
					xl[R] = ch.m > 1 ? ld(xb + (long long)R * S) : 0.0;
					xl[R + 1] = ch.m > 2 ? ld(xb + (long long)(R + 1) * S) : 0.0;
					if (c == 0) for (int q = 0; q < R; q++) xc[q] = t0 == 0 ? xl[q] : (t0 == 1 ? xl[q + 1] : xl[q + 2]);
					for (int q = 0; q < R; q++) sum[q] = std::fma(ch.v[0], xl[q], sum[q]);
					if (ch.m > 1) for (int q = 0; q < R; q++) sum[q] = std::fma(ch.v[1], xl[q + 1], sum[q]);
					if (ch.m > 2) for (int q = 0; q < R; q++) sum[q] = std::fma(ch.v[2], xl[q + 2], sum[q]);
				}
				if (t0 < 0) for (int q = 0; q < R; q++) xc[q] = ld(row0 + (long long)q * S);
				for (int q = 0; q < R; q++)
				{
					const long long row = row0 + (long long)q * S;
					y[(size_t)row] = sum[q]; written[(size_t)row]++; xc_seen[(size_t)row] = xc[q];
				}
			}
			else if (masked && p >= 0)
			{
				const int nch = info[(size_t)p] & 255;
				double sum[R];
				for (int q = 0; q < R; q++) sum[q] = 0.0;
				for (int c = 0; c < nch; c++)
				{
					const PatChainH& ch = chains[(size_t)p][(size_t)c];
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
					for (int q = 0; q < R; q++)
					{
						if (b[q] & 1u) sum[q] = std::fma(ch.v[0], xl[q], sum[q]);
						if (b[q] & 2u) sum[q] = std::fma(ch.v[1], xl[q + 1], sum[q]);
						if (b[q] & 4u) sum[q] = std::fma(ch.v[2], xl[q + 2], sum[q]);
					}
				}
				for (int q = 0; q < R; q++)
				{
					const long long row = row0 + (long long)q * S;
					if (row < n) { masked_rows++; y[(size_t)row] = sum[q]; written[(size_t)row]++; xc_seen[(size_t)row] = ld(row); }
				}
			}
			else
			{
				for (int q = 0; q < R; q++)
				{
					const long long row = row0 + (long long)q * S;
					if (row >= n) break;
					slow_rows++;
					const int pq = pat[(size_t)row];
					const int nch = info[(size_t)pq] & 255;
					double sum = 0.0;
					for (int c = 0; c < nch; c++)
					{
						const PatChainH& ch = chains[(size_t)pq][(size_t)c];
						for (int t = 0; t < ch.m; t++) sum = std::fma(ch.v[t], ld(row + ch.off + (long long)t * S), sum);
					}
					y[(size_t)row] = sum; written[(size_t)row]++; xc_seen[(size_t)row] = ld(row);
				}
			}
		}
	}
	// 3. against the CSR product
	double err = 0.0, scale = 0.0;
	int once = 0, xc_bad = 0;
	for (int r = 0; r < n; r++)
	{
		double s = 0.0;
		for (int k = A.rp[(size_t)r]; k < A.rp[(size_t)r + 1]; k++) s = std::fma(A.v[(size_t)k], x[(size_t)A.ci[(size_t)k]], s);
		err = std::max(err, std::fabs(s - y[(size_t)r])); scale = std::max(scale, std::fabs(s));
		once += written[(size_t)r] == 1;
		xc_bad += xc_seen[(size_t)r] != x[(size_t)r];
	}
	const long long nnz = A.rp[(size_t)n];
	std::printf("%-28s n=%-8d patterns=%-3zu S=%-6d chains(longest)=%-2zu items=%d uniform=%.3f masked_rows=%.4f slow_rows=%.4f loads/nnz=%.3f err=%.2e\n",
		name.c_str(), n, rows.size(), S, chains[longest].size(), n_items, (double)fast_items / std::max(1, n_items),
		(double)masked_rows / n, (double)slow_rows / n, (double)loads / (double)nnz, err / scale);
	int fail = 0;
	if (once != n) { std::printf("  rows written exactly once: %d of %d\n", once, n); fail = 1; }
	if (bad) { std::printf("  %d loads outside [0, n_cols)\n", bad); fail = 1; }
	if (xc_bad) { std::printf("  %d rows handed a wrong x[row] to the epilogue\n", xc_bad); fail = 1; }
	if (!(err / scale < 1e-14)) { std::printf("  y differs from the CSR product\n"); fail = 1; }
	if (expect_stride && S != expect_stride) { std::printf("  expected stride %d\n", expect_stride); fail = 1; }
	if (expect_chains && chains[longest].size() != expect_chains) { std::printf("  expected %zu chains\n", expect_chains); fail = 1; }
	return fail;
}

int main()
{
	int fail = 0;
	fail |= check("27pt 64^3", stencil(64, 64, 64, true, 0, 64), 64, 9);
	fail |= check("27pt 44^3", stencil(44, 44, 44, true, 0, 44), 44, 9);
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
