// ic0_host.h — zero-fill incomplete Cholesky (IC(0)) of a symmetric matrix, host side, and the level analysis for the
// triangular solves that apply it on the GPU.  The factorisation is the reference's sequential algorithm
// (preconditioner.cpp:33-160 real, preconditioner_cuda.cu:40-270 complex; L L^T with UNconjugated transposes for complex
// symmetric matrices), restated on a row-pointer view of the lower triangle with the very same operations in the very same
// order — each entry (i, c): subtract L(i,k) L(c,k) for EVERY earlier entry k of row i (a zero where row c has no entry at k,
// as the reference's scratch row does), divide by L(c,c), accumulate the square into the diagonal's sum — so the factor is
// bit-identical to the reference's (tests/test_oracle.py pins it against the compiled reference).
// Pure C++ (cuComplex.h only for the complex arithmetic the reference itself uses on the host): included by the CUDA library
// and by liblcg_dropin.cpp (lcg_incomplete_Cholesky_*_coo, clcg_incomplete_Cholesky_cuda_*).
#pragma once
#include <cmath>
#include <complex>
#include <vector>
#include <cuComplex.h>

namespace lcgb200 {

struct IcReal {
	typedef double T;
	static T zero() { return 0.0; }
	static T sub(T a, T b) { return a - b; }
	static T add(T a, T b) { return a + b; }
	static T mul(T a, T b) { return a * b; }
	static T div(T a, T b) { return a / b; }
	static T root(T a) { return std::sqrt(a); }
};
struct IcCplx {	// cuCmul / cuCdiv / std::sqrt(std::complex) as preconditioner_cuda.cu:156-270 and lcg_complex_cuda.cu:232-238 use them
	typedef cuDoubleComplex T;
	static T zero() { return make_cuDoubleComplex(0.0, 0.0); }
	static T sub(T a, T b) { return make_cuDoubleComplex(a.x - b.x, a.y - b.y); }
	static T add(T a, T b) { return make_cuDoubleComplex(a.x + b.x, a.y + b.y); }
	static T mul(T a, T b) { return cuCmul(a, b); }
	static T div(T a, T b) { return cuCdiv(a, b); }
	static T root(T a) { const std::complex<double> c = std::sqrt(std::complex<double>(a.x, a.y)); return make_cuDoubleComplex(c.real(), c.imag()); }
};
struct IcCplxF {	// preconditioner_cuda.cu:40-154
	typedef cuComplex T;
	static T zero() { return make_cuComplex(0.f, 0.f); }
	static T sub(T a, T b) { return make_cuComplex(a.x - b.x, a.y - b.y); }
	static T add(T a, T b) { return make_cuComplex(a.x + b.x, a.y + b.y); }
	static T mul(T a, T b) { return cuCmulf(a, b); }
	static T div(T a, T b) { return cuCdivf(a, b); }
	static T root(T a) { const std::complex<float> c = std::sqrt(std::complex<float>(a.x, a.y)); return make_cuComplex(c.real(), c.imag()); }
};

// In-place IC(0) on the lower triangle: rows 0..n-1, row i = entries [rp[i], rp[i+1]) with ascending columns <= i, the
// diagonal last.  Returns false when a row has no diagonal entry (the reference would read out of bounds there).
template <class M>
bool ic0_lower(int n, const int* rp, const int* col, typename M::T* val)
{
	typedef typename M::T T;
	std::vector<T> diag((size_t)n, M::zero()), tmp((size_t)n, M::zero());
	for (int i = 0; i < n; i++)
	{
		const int b = rp[i], e = rp[i + 1];
		if (e <= b || col[e - 1] != i) return false;
		T dia_sum = M::zero();
		for (int p = b; p < e; p++)
		{
			const int c = col[p];
			if (c < i)
			{
				const int cb = rp[c], ce = rp[c + 1];
				for (int q = cb; q < ce && col[q] < c; q++) tmp[(size_t)col[q]] = val[q];          // row c of L, left of its diagonal
				for (int q = b; q < p; q++) val[p] = M::sub(val[p], M::mul(val[q], tmp[(size_t)col[q]]));
				val[p] = M::div(val[p], diag[(size_t)c]);
				dia_sum = M::add(dia_sum, M::mul(val[p], val[p]));
				for (int q = cb; q < ce && col[q] < c; q++) tmp[(size_t)col[q]] = M::zero();
			}
			else
			{
				val[p] = M::root(M::sub(val[p], dia_sum));
				diag[(size_t)i] = val[p];
				dia_sum = M::zero();
			}
		}
	}
	return true;
}

// Levels of a triangular solve: lower (forward) — level[i] = 1 + max level of the columns left of the diagonal in row i;
// `order` lists the rows level by level, every level padded with -1 to a multiple of `pad` positions (a warp of the solve
// kernel then never holds two rows that depend on each other).  Returns the number of levels.
inline int level_order(int n, const int* rp, const int* col, bool upper, int pad, std::vector<int>& order)
{
	std::vector<int> level((size_t)n, 0);
	int n_levels = 0;
	for (int s = 0; s < n; s++)
	{
		const int i = upper ? n - 1 - s : s;
		int lv = 0;
		for (int k = rp[i]; k < rp[i + 1]; k++)
		{
			const int c = col[k];
			if (c != i && level[(size_t)c] + 1 > lv) lv = level[(size_t)c] + 1;
		}
		level[(size_t)i] = lv;
		if (lv + 1 > n_levels) n_levels = lv + 1;
	}
	std::vector<int> count((size_t)n_levels, 0), start((size_t)n_levels + 1, 0);
	for (int i = 0; i < n; i++) count[(size_t)level[(size_t)i]]++;
	for (int l = 0; l < n_levels; l++) start[(size_t)l + 1] = start[(size_t)l] + (count[(size_t)l] + pad - 1) / pad * pad;
	order.assign((size_t)start[(size_t)n_levels], -1);
	std::vector<int> fill(start.begin(), start.end() - 1);
	for (int i = 0; i < n; i++) order[(size_t)fill[(size_t)level[(size_t)i]]++] = i;
	return n_levels;
}

}  // namespace lcgb200
