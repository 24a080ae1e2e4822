// liblcg_dropin.cpp — liblcg_dropin.so: liblcg's C++ symbols (liblcg_abi.h), defined out of line and forwarding to the C ABI
// of liblcgb200.so.  A program compiled against the REFERENCE's own headers (lcg.h, clcg.h, lcg_cuda.h, clcg_cuda.h,
// solver.h, solver_cuda.h, util.h, algebra.h, lcg_complex.h) links against this library instead of liblcg.so — or is
// re-pointed at it at load time — and runs on the B200-native engine: its Ax/Mx/progress callbacks are honoured on the
// generic path, and passing the exported sentinels (lcgb200_csr_ax & co.) selects the fused built-in CSR operator.
//
// The small host helpers of algebra.h / lcg_complex.h are provided because every caller of the solvers uses them around
// the call (vector allocation, fill, dot); they are host-side conveniences, not part of the hot path.
#include "liblcg_abi.h"
#include "../../../include/lcgb200.h"
#include "../ic0_host.h"
#include <vector>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <stdexcept>

static_assert(sizeof(lcg_para) == sizeof(lcgb200_para) && sizeof(clcg_para) == sizeof(lcgb200_cpara), "parameter blocks pass through the C ABI as they are");

namespace {
inline const lcgb200_para* P(const lcg_para* p) { return reinterpret_cast<const lcgb200_para*>(p); }
inline const lcgb200_cpara* P(const clcg_para* p) { return reinterpret_cast<const lcgb200_cpara*>(p); }
const lcg_para kDef = {0, 1e-6, 0, 1e-6, 1.0, 0.95, 0.9, 10};   // util.h:153
const clcg_para kDefC = {0, 1e-6, 0};                           // util.h:278
}

// ------------------------------------------------------------------------------------------------ util.h
lcg_para lcg_default_parameters() { return kDef; }
clcg_para clcg_default_parameters() { return kDefC; }

lcg_solver_enum lcg_select_solver(std::string name)
{	// util.cpp:39-51: unknown names fall back to CGS
	static const char* names[] = {"LCG_CG", "LCG_PCG", "LCG_CGS", "LCG_BICGSTAB", "LCG_BICGSTAB2", "LCG_PG", "LCG_SPG"};
	for (int i = 0; i < 7; i++) if (name == names[i]) return static_cast<lcg_solver_enum>(i);
	return LCG_CGS;
}
clcg_solver_enum clcg_select_solver(std::string name)
{	// util.cpp:157-166 knows four names
	if (name == "CLCG_BICG") return CLCG_BICG;
	if (name == "CLCG_BICG_SYM") return CLCG_BICG_SYM;
	if (name == "CLCG_TFQMR") return CLCG_TFQMR;
	return CLCG_CGS;
}

namespace {
const char* code_text(int code, bool cplx)
{
	if (code == 0) return "The iteration reached convergence.";
	if (code == 1) return "The iteration was stopped by the progress callback.";
	if (code == 2) return "The initial solution is already optimized.";
	if (cplx)
		switch (code)
		{
			case CLCG_INVILAD_VARIABLE_SIZE: return "The variable size is not positive.";
			case CLCG_INVILAD_MAX_ITERATIONS: return "The maximal iteration count is negative.";
			case CLCG_INVILAD_EPSILON: return "The epsilon is not in (0,1).";
			case CLCG_REACHED_MAX_ITERATIONS: return "The iteration reached the maximal limit.";
			case CLCG_NAN_VALUE: return "The model values are NaN.";
			case CLCG_INVALID_POINTER: return "Invalid pointer.";
			case CLCG_SIZE_NOT_MATCH: return "The sizes of the operator and the vectors do not match.";
			case CLCG_UNKNOWN_SOLVER: return "Unknown solver.";
			default: return "Unknown error.";
		}
	switch (code)
	{
		case LCG_INVILAD_VARIABLE_SIZE: return "The variable size is not positive.";
		case LCG_INVILAD_MAX_ITERATIONS: return "The maximal iteration count is negative.";
		case LCG_INVILAD_EPSILON: return "The epsilon is not in (0,1).";
		case LCG_INVILAD_RESTART_EPSILON: return "The restart epsilon is not positive.";
		case LCG_REACHED_MAX_ITERATIONS: return "The iteration reached the maximal limit.";
		case LCG_NULL_PRECONDITION_MATRIX: return "The preconditioner is missing.";
		case LCG_NAN_VALUE: return "The model values are NaN.";
		case LCG_INVALID_POINTER: return "Invalid pointer.";
		case LCG_INVALID_LAMBDA: return "Invalid range for lambda (step).";
		case LCG_INVALID_SIGMA: return "Invalid range for sigma.";
		case LCG_INVALID_BETA: return "Invalid range for beta.";
		case LCG_INVALID_MAXIM: return "Invalid range for maxi_m.";
		case LCG_SIZE_NOT_MATCH: return "The sizes of the operator and the vectors do not match.";
		default: return "Unknown error.";
	}
}
}  // namespace

void lcg_error_str(int er_index, bool er_throw)
{	// util.cpp:53-148: one line on stderr; with er_throw a negative code raises (the only thrower of the library, util.cpp:120)
	const char* text = code_text(er_index, false);
	if (er_throw && er_index < 0) throw std::runtime_error(std::string("[LibLCG] ") + text);
	std::fprintf(stderr, "%s %s\n", er_index >= 0 ? "Success!" : (er_index == LCG_REACHED_MAX_ITERATIONS ? "Warning!" : "Fail!"), text);
}
void clcg_error_str(int er_index, bool er_throw)
{
	const char* text = code_text(er_index, true);
	if (er_throw && er_index < 0) throw std::runtime_error(std::string("[LibLCG] ") + text);
	std::fprintf(stderr, "%s %s\n", er_index >= 0 ? "Success!" : "Fail!", text);
}

// ------------------------------------------------------------------------------------ algebra.h host helpers
lcg_float lcg_abs(lcg_float a) { return a >= 0.0 ? a : -a; }
lcg_float lcg_max(lcg_float a, lcg_float b) { return a >= b ? a : b; }
lcg_float lcg_min(lcg_float a, lcg_float b) { return a <= b ? a : b; }
lcg_float lcg_set2box(lcg_float low, lcg_float hig, lcg_float a, bool low_bound, bool hig_bound)
{	// algebra.cpp:50-58: closed bounds clamp onto the bound, open bounds stop a hair inside
	if (hig_bound && a >= hig) return hig;
	if (!hig_bound && a >= hig) return hig - 1e-16;
	if (low_bound && a <= low) return low;
	if (!low_bound && a <= low) return low + 1e-16;
	return a;
}
lcg_float* lcg_malloc(int n) { return new lcg_float[n]; }
lcg_float** lcg_malloc(int m, int n)
{
	lcg_float** x = new lcg_float*[m];
	for (int i = 0; i < m; i++) x[i] = new lcg_float[n];
	return x;
}
void lcg_free(lcg_float* x) { delete[] x; }
void lcg_free(lcg_float** x, int m)
{
	if (!x) return;
	for (int i = 0; i < m; i++) delete[] x[i];
	delete[] x;
}
void lcg_vecset(lcg_float* a, lcg_float b, int size) { for (int i = 0; i < size; i++) a[i] = b; }
void lcg_vecset(lcg_float** a, lcg_float b, int m, int n) { for (int i = 0; i < m; i++) lcg_vecset(a[i], b, n); }
void lcg_vecrnd(lcg_float* a, lcg_float l, lcg_float h, int size)
{	// algebra.cpp: seeded from the clock on every call, uniform in [l, h]
	srand((unsigned)time(nullptr));
	for (int i = 0; i < size; i++) a[i] = (h - l) * rand() / RAND_MAX + l;
}
void lcg_vecrnd(lcg_float** a, lcg_float l, lcg_float h, int m, int n)
{
	srand((unsigned)time(nullptr));
	for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) a[i][j] = (h - l) * rand() / RAND_MAX + l;
}
double lcg_squaredl2norm(lcg_float* a, int n) { double s = 0.0; for (int i = 0; i < n; i++) s += a[i] * a[i]; return s; }
void lcg_dot(lcg_float& ret, const lcg_float* a, const lcg_float* b, int size)
{	// algebra.cpp:154-163: left-to-right serial sum
	ret = 0.0;
	for (int i = 0; i < size; i++) ret += a[i] * b[i];
}
void lcg_matvec(lcg_float** A, const lcg_float* x, lcg_float* Ax, int m_size, int n_size, lcg_matrix_e layout)
{
	if (layout == MatNormal)
		for (int i = 0; i < m_size; i++) { lcg_float s = 0.0; for (int j = 0; j < n_size; j++) s += A[i][j] * x[j]; Ax[i] = s; }
	else
		for (int j = 0; j < n_size; j++) { lcg_float s = 0.0; for (int i = 0; i < m_size; i++) s += A[i][j] * x[i]; Ax[j] = s; }
}
void lcg_matvec_coo(const int* row, const int* col, const lcg_float* Mat, const lcg_float* V, lcg_float* p, int M, int N, int nz_size, bool pre_position)
{	// p = A V, or p = A^T V when pre_position
	const int out = pre_position ? N : M;
	for (int i = 0; i < out; i++) p[i] = 0.0;
	if (pre_position) for (int k = 0; k < nz_size; k++) p[col[k]] += Mat[k] * V[row[k]];
	else for (int k = 0; k < nz_size; k++) p[row[k]] += Mat[k] * V[col[k]];
}

// -------------------------------------------------------------------------------- lcg_complex.h host helpers
lcg_complex* clcg_malloc(int n) { return new lcg_complex[n]; }
lcg_complex** clcg_malloc(int m, int n)
{
	lcg_complex** x = new lcg_complex*[m];
	for (int i = 0; i < m; i++) x[i] = new lcg_complex[n];
	return x;
}
void clcg_free(lcg_complex* x) { delete[] x; }
void clcg_free(lcg_complex** x, int m)
{
	if (!x) return;
	for (int i = 0; i < m; i++) delete[] x[i];
	delete[] x;
}
void clcg_vecset(lcg_complex* a, lcg_complex b, int size) { for (int i = 0; i < size; i++) a[i] = b; }
void clcg_vecset(lcg_complex** a, lcg_complex b, int m, int n) { for (int i = 0; i < m; i++) clcg_vecset(a[i], b, n); }
void clcg_set(lcg_complex* a, lcg_float r, lcg_float i) { *a = lcg_complex(r, i); }
lcg_float clcg_square(const lcg_complex* a) { return std::norm(*a); }
lcg_float clcg_module(const lcg_complex* a) { return std::sqrt(std::norm(*a)); }
lcg_complex clcg_conjugate(const lcg_complex* a) { return std::conj(*a); }
void clcg_vecrnd(lcg_complex* a, lcg_complex l, lcg_complex h, int size)
{	// lcg_complex.cpp:118-127: one rand() per component, real first
	srand((unsigned)time(nullptr));
	for (int i = 0; i < size; i++)
	{
		const lcg_float re = (h.real() - l.real()) * rand() / RAND_MAX + l.real();
		const lcg_float im = (h.imag() - l.imag()) * rand() / RAND_MAX + l.imag();
		a[i] = lcg_complex(re, im);
	}
}
void clcg_vecrnd(lcg_complex** a, lcg_complex l, lcg_complex h, int m, int n)
{
	srand((unsigned)time(nullptr));
	for (int i = 0; i < m; i++)
		for (int j = 0; j < n; j++)
		{
			const lcg_float re = (h.real() - l.real()) * rand() / RAND_MAX + l.real();
			const lcg_float im = (h.imag() - l.imag()) * rand() / RAND_MAX + l.imag();
			a[i][j] = lcg_complex(re, im);
		}
}
void clcg_dot(lcg_complex& ret, const lcg_complex* a, const lcg_complex* b, int size)
{	// unconjugated (lcg_complex.cpp:143-153)
	ret = lcg_complex(0.0, 0.0);
	for (int i = 0; i < size; i++) ret += a[i] * b[i];
}
void clcg_inner(lcg_complex& ret, const lcg_complex* a, const lcg_complex* b, int size)
{	// conjugate of the FIRST argument (lcg_complex.cpp:155-167)
	ret = lcg_complex(0.0, 0.0);
	for (int i = 0; i < size; i++) ret += std::conj(a[i]) * b[i];
}
void clcg_matvec(lcg_complex** A, const lcg_complex* x, lcg_complex* Ax, int m_size, int n_size, lcg_matrix_e layout, clcg_complex_e conjugate)
{
	const bool cj = conjugate == Conjugate;
	if (layout == MatNormal)
		for (int i = 0; i < m_size; i++) { lcg_complex s(0.0, 0.0); for (int j = 0; j < n_size; j++) s += (cj ? std::conj(A[i][j]) : A[i][j]) * x[j]; Ax[i] = s; }
	else
		for (int j = 0; j < n_size; j++) { lcg_complex s(0.0, 0.0); for (int i = 0; i < m_size; i++) s += (cj ? std::conj(A[i][j]) : A[i][j]) * x[i]; Ax[j] = s; }
}

// ------------------------------------------------------------------------ algebra_cuda.h / lcg_complex_cuda.h
// Device helpers: forwarded to the element-wise kernels of liblcgb200.so on the default stream, like the reference's
// launches (algebra_cuda.cu:79-110, lcg_complex_cuda.cu:294-356); the block-size argument has no meaning here.
void lcg_set2box_cuda(const lcg_float* low, const lcg_float* hig, lcg_float* a, int n, bool, bool) { lcgb200_set2box(low, hig, a, n, nullptr); }
void lcg_smDcsr_get_diagonal(const int* A_ptr, const int* A_col, const lcg_float* A_val, const int A_len, lcg_float* A_diag, int)
{
	lcgb200_diagonal_of_csr(LCGB200_REAL, A_ptr, A_col, A_val, A_len, A_diag, nullptr);
}
void lcg_vecMvecD_element_wise(const lcg_float* a, const lcg_float* b, lcg_float* c, int n, int) { lcgb200_vec_elementwise(0, LCGB200_REAL, a, b, c, n, nullptr); }
void lcg_vecDvecD_element_wise(const lcg_float* a, const lcg_float* b, lcg_float* c, int n, int) { lcgb200_vec_elementwise(1, LCGB200_REAL, a, b, c, n, nullptr); }
void clcg_smCcsr_get_diagonal(const int* A_ptr, const int* A_col, const cuComplex* A_val, const int A_len, cuComplex* A_diag, int)
{
	lcgb200_diagonal_of_csr(LCGB200_COMPLEX_FLOAT, A_ptr, A_col, A_val, A_len, A_diag, nullptr);
}
void clcg_smZcsr_get_diagonal(const int* A_ptr, const int* A_col, const cuDoubleComplex* A_val, const int A_len, cuDoubleComplex* A_diag, int)
{
	lcgb200_diagonal_of_csr(LCGB200_COMPLEX, A_ptr, A_col, A_val, A_len, A_diag, nullptr);
}
void clcg_vecMvecC_element_wise(const cuComplex* a, const cuComplex* b, cuComplex* c, int n, int) { lcgb200_vec_elementwise(0, LCGB200_COMPLEX_FLOAT, a, b, c, n, nullptr); }
void clcg_vecMvecZ_element_wise(const cuDoubleComplex* a, const cuDoubleComplex* b, cuDoubleComplex* c, int n, int) { lcgb200_vec_elementwise(0, LCGB200_COMPLEX, a, b, c, n, nullptr); }
void clcg_vecDvecC_element_wise(const cuComplex* a, const cuComplex* b, cuComplex* c, int n, int) { lcgb200_vec_elementwise(1, LCGB200_COMPLEX_FLOAT, a, b, c, n, nullptr); }
void clcg_vecDvecZ_element_wise(const cuDoubleComplex* a, const cuDoubleComplex* b, cuDoubleComplex* c, int n, int) { lcgb200_vec_elementwise(1, LCGB200_COMPLEX, a, b, c, n, nullptr); }
void clcg_vecC_conjugate(const cuComplex* a, cuComplex* ca, int n, int) { lcgb200_vec_elementwise(2, LCGB200_COMPLEX_FLOAT, a, nullptr, ca, n, nullptr); }
void clcg_vecZ_conjugate(const cuDoubleComplex* a, cuDoubleComplex* ca, int n, int) { lcgb200_vec_elementwise(2, LCGB200_COMPLEX, a, nullptr, ca, n, nullptr); }

// Host helpers around cuComplex values (lcg_complex_cuda.cu:133-238)
lcg_complex cuda2lcg_complex(cuDoubleComplex a) { return lcg_complex(a.x, a.y); }
cuDoubleComplex lcg2cuda_complex(lcg_complex a) { return make_cuDoubleComplex(a.real(), a.imag()); }
cuDoubleComplex* clcg_malloc_cuda(size_t n) { return new cuDoubleComplex[n]; }   // host memory, despite the name (lcg_complex_cuda.cu:155-159)
void clcg_free_cuda(cuDoubleComplex* x) { delete[] x; }
void clcg_vecset_cuda(cuDoubleComplex* a, cuDoubleComplex b, size_t size) { for (size_t i = 0; i < size; i++) a[i] = b; }
cuComplex clcg_Cscale(float s, cuComplex a) { return make_cuComplex(s * a.x, s * a.y); }
cuComplex clcg_Csum(cuComplex a, cuComplex b) { return make_cuComplex(a.x + b.x, a.y + b.y); }
cuComplex clcg_Cdiff(cuComplex a, cuComplex b) { return make_cuComplex(a.x - b.x, a.y - b.y); }
cuComplex clcg_Csqrt(cuComplex a) { const std::complex<float> c = std::sqrt(std::complex<float>(a.x, a.y)); return make_cuComplex(c.real(), c.imag()); }
cuDoubleComplex clcg_Zscale(lcg_float s, cuDoubleComplex a) { return make_cuDoubleComplex(s * a.x, s * a.y); }
cuDoubleComplex clcg_Zsum(cuDoubleComplex a, cuDoubleComplex b) { return make_cuDoubleComplex(a.x + b.x, a.y + b.y); }
cuDoubleComplex clcg_Zdiff(cuDoubleComplex a, cuDoubleComplex b) { return make_cuDoubleComplex(a.x - b.x, a.y - b.y); }
cuDoubleComplex clcg_Zsqrt(cuDoubleComplex a) { const std::complex<lcg_float> c = std::sqrt(std::complex<lcg_float>(a.x, a.y)); return make_cuDoubleComplex(c.real(), c.imag()); }

// Row-sorted COO -> column-sorted COO with rows and columns exchanged, i.e. the transpose in row-sorted COO
// (lcg_complex_cuda.cu:240-292 sorts through a std::map keyed N*col + row: a repeated (row, col) keeps its LAST value and the
// output is shorter than nz by the number of repeats).  Here: a stable sort of the entry indices by that key.
template <class V>
static void coo_row2col(const int* A_row, const int* A_col, const V* A, int N, int nz, int* Ac_row, int* Ac_col, V* Ac_val)
{
	std::vector<size_t> order((size_t)std::max(nz, 0));
	for (size_t i = 0; i < order.size(); i++) order[i] = i;
	auto key = [&](size_t i) { return (size_t)N * (size_t)A_col[i] + (size_t)A_row[i]; };
	std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return key(a) < key(b); });
	size_t out = 0;
	for (size_t k = 0; k < order.size(); k++)
	{
		if (k + 1 < order.size() && key(order[k + 1]) == key(order[k])) continue;   // a later entry with the same key overwrites this one
		Ac_row[out] = A_col[order[k]]; Ac_col[out] = A_row[order[k]]; Ac_val[out] = A[order[k]];
		out++;
	}
}
void clcg_smCcoo_row2col(const int* A_row, const int* A_col, const cuComplex* A, int N, int nz, int* Ac_row, int* Ac_col, cuComplex* Ac_val)
{
	coo_row2col(A_row, A_col, A, N, nz, Ac_row, Ac_col, Ac_val);
}
void clcg_smZcoo_row2col(const int* A_row, const int* A_col, const cuDoubleComplex* A, int N, int nz, int* Ac_row, int* Ac_col, cuDoubleComplex* Ac_val)
{
	coo_row2col(A_row, A_col, A, N, nz, Ac_row, Ac_col, Ac_val);
}

// ------------------------------------------------------------------- preconditioner.h / preconditioner_cuda.h
// IC(0) on row-sorted COO input: the lower triangle is copied out, factorised by the shared host routine (ic0_host.h: the
// reference's operations in the reference's order, bit-identical factor) and, for the *_full variants, mirrored into the
// upper positions of the output (which keeps the pattern of A).
namespace {
template <class M>
void ic0_half_coo(const int* row, const int* col, const typename M::T* val, int N, int nz, int* ic_row, int* ic_col, typename M::T* ic_val)
{
	std::vector<int> rp((size_t)N + 1, 0);
	int j = 0;
	for (int i = 0; i < nz; i++)
		if (row[i] >= col[i]) { ic_row[j] = row[i]; ic_col[j] = col[i]; ic_val[j] = val[i]; rp[(size_t)row[i] + 1]++; j++; }
	for (int i = 0; i < N; i++) rp[(size_t)i + 1] += rp[(size_t)i];
	lcgb200::ic0_lower<M>(N, rp.data(), ic_col, ic_val);
}
template <class M>
void ic0_full_coo(const int* row, const int* col, const typename M::T* val, int N, int nz, int* ic_row, int* ic_col, typename M::T* ic_val)
{
	int lnz = 0;
	for (int i = 0; i < nz; i++) lnz += row[i] >= col[i];
	std::vector<int> lr((size_t)lnz), lc((size_t)lnz); std::vector<typename M::T> lv((size_t)lnz);
	ic0_half_coo<M>(row, col, val, N, nz, lr.data(), lc.data(), lv.data());
	// position of every lower entry, to mirror L(i,c) into (c,i)
	std::vector<int> rp((size_t)N + 1, 0);
	for (int k = 0; k < lnz; k++) rp[(size_t)lr[(size_t)k] + 1]++;
	for (int i = 0; i < N; i++) rp[(size_t)i + 1] += rp[(size_t)i];
	int l = 0;
	for (int i = 0; i < nz; i++)
	{
		ic_row[i] = row[i]; ic_col[i] = col[i];
		if (row[i] >= col[i]) ic_val[i] = lv[(size_t)l++];
		else
		{	// upper entry (r, c), r < c: the value of L(c, r) if the lower triangle holds it
			ic_val[i] = val[i];
			for (int k = rp[(size_t)col[i]]; k < rp[(size_t)col[i] + 1]; k++) if (lc[(size_t)k] == row[i]) { ic_val[i] = lv[(size_t)k]; break; }
		}
	}
}
}  // namespace

void lcg_incomplete_Cholesky_half_buffsize_coo(const int* row, const int* col, int nz_size, int* lnz_size)
{
	int c = 0;
	for (int i = 0; i < nz_size; i++) c += row[i] >= col[i];
	*lnz_size = c;
}
void lcg_incomplete_Cholesky_half_coo(const int* row, const int* col, const lcg_float* val, int N, int nz_size, int, int* IC_row, int* IC_col, lcg_float* IC_val)
{
	ic0_half_coo<lcgb200::IcReal>(row, col, val, N, nz_size, IC_row, IC_col, IC_val);
}
void lcg_incomplete_Cholesky_full_coo(const int* row, const int* col, const lcg_float* val, int N, int nz_size, int* IC_row, int* IC_col, lcg_float* IC_val)
{
	ic0_full_coo<lcgb200::IcReal>(row, col, val, N, nz_size, IC_row, IC_col, IC_val);
}
void clcg_incomplete_Cholesky_cuda_half_buffsize(const int* row, const int* col, int nz_size, int* lnz_size)
{
	lcg_incomplete_Cholesky_half_buffsize_coo(row, col, nz_size, lnz_size);
}
void clcg_incomplete_Cholesky_cuda_half(const int* row, const int* col, const cuComplex* val, int N, int nz_size, int, int* IC_row, int* IC_col, cuComplex* IC_val)
{
	ic0_half_coo<lcgb200::IcCplxF>(row, col, val, N, nz_size, IC_row, IC_col, IC_val);
}
void clcg_incomplete_Cholesky_cuda_half(const int* row, const int* col, const cuDoubleComplex* val, int N, int nz_size, int, int* IC_row, int* IC_col,
	cuDoubleComplex* IC_val)
{
	ic0_half_coo<lcgb200::IcCplx>(row, col, val, N, nz_size, IC_row, IC_col, IC_val);
}
void clcg_incomplete_Cholesky_cuda_full(const int* row, const int* col, const cuDoubleComplex* val, int N, int nz_size, int* IC_row, int* IC_col,
	cuDoubleComplex* IC_val)
{
	ic0_full_coo<lcgb200::IcCplx>(row, col, val, N, nz_size, IC_row, IC_col, IC_val);
}
void lcg_solve_lower_triangle_coo(const int* row, const int* col, const lcg_float* L, const lcg_float* B, lcg_float* x, int N, int nz_size)
{	// forward substitution over row-sorted COO, the row's sum accumulated left to right (preconditioner.cpp:340-366)
	for (int i = 0; i < N; i++) x[i] = 0.0;
	int k = 0;
	for (int i = 0; i < N; i++)
	{
		double sum = 0.0;
		for (; k < nz_size && row[k] == i; k++)
		{
			if (col[k] < i) sum += L[k] * x[col[k]];
			else if (col[k] == i) { x[i] = (B[i] - sum) / L[k]; }
		}
	}
}
void lcg_solve_upper_triangle_coo(const int* row, const int* col, const lcg_float* U, const lcg_float* B, lcg_float* x, int N, int nz_size)
{	// backward substitution, the row's sum accumulated right to left (preconditioner.cpp:300-338)
	for (int i = 0; i < N; i++) x[i] = 0.0;
	int k = nz_size - 1;
	for (int i = N - 1; i >= 0; i--)
	{
		double sum = 0.0;
		for (; k >= 0 && row[k] == i; k--)
		{
			if (col[k] > i) sum += U[k] * x[col[k]];
			else if (col[k] == i) { x[i] = (B[i] - sum) / U[k]; }
		}
	}
}
bool lcg_full_rank_coo(const int* row, const int* col, const lcg_float* M, int N, int nz_size)
{
	int s = 0;
	for (int i = 0; i < nz_size; i++) s += (row[i] == col[i] && M[i] != 0.0);
	return s == N;
}

// --------------------------------------------------------------------------------------- host-callback solvers
int lcg_solver(lcg_axfunc_ptr Afp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size, const lcg_para* param, void* instance,
	lcg_solver_enum solver_id)
{
	return lcgb200_solver(Afp, reinterpret_cast<lcgb200_progress_ptr>(Pfp), m, B, n_size, P(param), instance, static_cast<int>(solver_id));
}
int lcg_solver_preconditioned(lcg_axfunc_ptr Afp, lcg_axfunc_ptr Mfp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size,
	const lcg_para* param, void* instance, lcg_solver_enum solver_id)
{
	return lcgb200_solver_preconditioned(Afp, Mfp, reinterpret_cast<lcgb200_progress_ptr>(Pfp), m, B, n_size, P(param), instance, static_cast<int>(solver_id));
}
int lcg_solver_constrained(lcg_axfunc_ptr Afp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const lcg_float* low, const lcg_float* hig,
	const int n_size, const lcg_para* param, void* instance, lcg_solver_enum solver_id)
{
	return lcgb200_solver_constrained(Afp, reinterpret_cast<lcgb200_progress_ptr>(Pfp), m, B, low, hig, n_size, P(param), instance, static_cast<int>(solver_id));
}
int lcg(lcg_axfunc_ptr Afp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size, const lcg_para* param, void* instance,
	lcg_float* Gk, lcg_float* Dk, lcg_float* ADk)
{
	return lcgb200_lcg(Afp, reinterpret_cast<lcgb200_progress_ptr>(Pfp), m, B, n_size, P(param), instance, Gk, Dk, ADk);
}
int lcgs(lcg_axfunc_ptr Afp, lcg_progress_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size, const lcg_para* param, void* instance,
	lcg_float* RK, lcg_float* R0T, lcg_float* PK, lcg_float* AX, lcg_float* UK, lcg_float* QK, lcg_float* WK)
{
	return lcgb200_lcgs(Afp, reinterpret_cast<lcgb200_progress_ptr>(Pfp), m, B, n_size, P(param), instance, RK, R0T, PK, AX, UK, QK, WK);
}
int clcg_solver(clcg_axfunc_ptr Afp, clcg_progress_ptr Pfp, lcg_complex* m, const lcg_complex* B, const int n_size, const clcg_para* param,
	void* instance, clcg_solver_enum solver_id)
{	// std::complex<double> is layout-compatible with interleaved (re, im); the enums travel as ints
	return lcgb200_csolver(reinterpret_cast<lcgb200_caxfunc_ptr>(Afp), reinterpret_cast<lcgb200_cprogress_ptr>(Pfp), m, B, n_size, P(param), instance,
		static_cast<int>(solver_id));
}

// ------------------------------------------------------------------------------------------- CUDA entry points
int lcg_solver_cuda(lcg_axfunc_cuda_ptr Afp, lcg_progress_cuda_ptr Pfp, lcg_float* m, const lcg_float* B, const int n_size, const int nz_size,
	const lcg_para* param, void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle, lcg_solver_enum solver_id)
{
	return lcgb200_solver_cuda(reinterpret_cast<lcgb200_axfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_progress_cuda_ptr>(Pfp), m, B, n_size, nz_size,
		P(param), instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle), reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}
int lcg_solver_preconditioned_cuda(lcg_axfunc_cuda_ptr Afp, lcg_axfunc_cuda_ptr Mfp, lcg_progress_cuda_ptr Pfp, lcg_float* m, const lcg_float* B,
	const int n_size, const int nz_size, const lcg_para* param, void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle,
	lcg_solver_enum solver_id)
{
	return lcgb200_solver_preconditioned_cuda(reinterpret_cast<lcgb200_axfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_axfunc_cuda_ptr>(Mfp),
		reinterpret_cast<lcgb200_progress_cuda_ptr>(Pfp), m, B, n_size, nz_size, P(param), instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle),
		reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}
int lcg_solver_constrained_cuda(lcg_axfunc_cuda_ptr Afp, lcg_progress_cuda_ptr Pfp, lcg_float* m, const lcg_float* B, const lcg_float* low,
	const lcg_float* hig, const int n_size, const int nz_size, const lcg_para* param, void* instance, cublasHandle_t cub_handle,
	cusparseHandle_t cus_handle, lcg_solver_enum solver_id)
{
	return lcgb200_solver_constrained_cuda(reinterpret_cast<lcgb200_axfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_progress_cuda_ptr>(Pfp), m, B, low, hig,
		n_size, nz_size, P(param), instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle), reinterpret_cast<lcgb200_cusparse_t>(cus_handle),
		static_cast<int>(solver_id));
}
int clcg_solver_cuda(clcg_axfunc_cuda_ptr Afp, clcg_progress_cuda_ptr Pfp, cuDoubleComplex* m, const cuDoubleComplex* B, const int n_size,
	const int nz_size, const clcg_para* param, void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle, clcg_solver_enum solver_id)
{
	return lcgb200_csolver_cuda(reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_cprogress_cuda_ptr>(Pfp), m, B, n_size, nz_size,
		P(param), instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle), reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}
int clcg_solver_preconditioned_cuda(clcg_axfunc_cuda_ptr Afp, clcg_axfunc_cuda_ptr Mfp, clcg_progress_cuda_ptr Pfp, cuDoubleComplex* m,
	const cuDoubleComplex* B, const int n_size, const int nz_size, const clcg_para* param, void* instance, cublasHandle_t cub_handle,
	cusparseHandle_t cus_handle, clcg_solver_enum solver_id)
{
	return lcgb200_csolver_preconditioned_cuda(reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Mfp),
		reinterpret_cast<lcgb200_cprogress_cuda_ptr>(Pfp), m, B, n_size, nz_size, P(param), instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle),
		reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}

// the cuComplex overloads (clcg_cudaf.h:81-83, 103-105)
int clcg_solver_cuda(clcg_axfunc_cudaf_ptr Afp, clcg_progress_cudaf_ptr Pfp, cuComplex* m, const cuComplex* B, const int n_size, const int nz_size,
	const clcg_para* param, void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle, clcg_solver_enum solver_id)
{
	return lcgb200_csolver_cudaf(reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_cprogress_cudaf_ptr>(Pfp), m, B, n_size, nz_size,
		P(param), instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle), reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}
int clcg_solver_preconditioned_cuda(clcg_axfunc_cudaf_ptr Afp, clcg_axfunc_cudaf_ptr Mfp, clcg_progress_cudaf_ptr Pfp, cuComplex* m, const cuComplex* B,
	const int n_size, const int nz_size, const clcg_para* param, void* instance, cublasHandle_t cub_handle, cusparseHandle_t cus_handle,
	clcg_solver_enum solver_id)
{
	return lcgb200_csolver_preconditioned_cudaf(reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Afp), reinterpret_cast<lcgb200_caxfunc_cuda_ptr>(Mfp),
		reinterpret_cast<lcgb200_cprogress_cudaf_ptr>(Pfp), m, B, n_size, nz_size, P(param), instance, reinterpret_cast<lcgb200_cublas_t>(cub_handle),
		reinterpret_cast<lcgb200_cusparse_t>(cus_handle), static_cast<int>(solver_id));
}

// ------------------------------------------------------------------------------------------- class wrappers
// solver.cpp:29-283 / solver_cuda.cu:29-414.  Minimize* print the solver's name, run, print the elapsed time and report the
// return code through lcg_error_str / clcg_error_str (which throws for a negative code when er_throw is set).
namespace {
const char* real_name(int id) { static const char* n[] = {"CG", "PCG", "CGS", "BICGSTAB", "BICGSTAB2", "PG", "SPG"}; return (id >= 0 && id < 7) ? n[id] : "Unknown"; }
const char* cplx_name(int id) { static const char* n[] = {"BICG", "BICG_SYM", "CGS", "BICGSTAB", "TFQMR", "PCG", "PBICG"}; return (id >= 0 && id < 7) ? n[id] : "Unknown"; }

int monitor(unsigned int inter, double converge, double epsilon, int k)
{	// the default Progress of every wrapper class (solver.cpp:36-51)
	if ((inter > 0 && k % inter == 0) || converge <= epsilon) std::clog << "\rIteration-times: " << k << "\tconvergence: " << converge;
	return 0;
}

template <class Call>
void minimize(bool silent, bool verbose, bool er_throw, bool cplx, const char* name, Call&& call)
{
	auto report = [&](int ret, bool thr) { if (cplx) clcg_error_str(ret, thr); else lcg_error_str(ret, thr); };
	if (silent)
	{	// solver.cpp:77-82: no monitor, and an error always raises
		const int ret = call(false);
		if (ret < 0) report(ret, true);
		return;
	}
	const auto t0 = std::chrono::steady_clock::now();
	const int ret = call(true);
	const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	if (!er_throw) std::clog << std::endl << "Solver: " << name << ". Time cost: " << ms << " ms" << std::endl;
	if (verbose || ret < 0) report(ret, er_throw);
}
}  // namespace

LCG_Solver::LCG_Solver() : param_(kDef), inter_(1), silent_(false) {}
int LCG_Solver::Progress(const lcg_float*, const lcg_float converge, const lcg_para* param, const int, const int k) { return monitor(inter_, converge, param->epsilon, k); }
void LCG_Solver::silent() { silent_ = true; }
void LCG_Solver::set_report_interval(unsigned int inter) { inter_ = inter; }
void LCG_Solver::set_lcg_parameter(const lcg_para& in_param) { param_ = in_param; }

namespace {
// the trampolines the reference keeps inline in its header (solver.h:52-75): the library needs its own to pass as callbacks
struct LcgTramp {
	static void ax(void* inst, const lcg_float* a, lcg_float* b, const int n) { static_cast<LCG_Solver*>(inst)->AxProduct(a, b, n); }
	static void mx(void* inst, const lcg_float* a, lcg_float* b, const int n) { static_cast<LCG_Solver*>(inst)->MxProduct(a, b, n); }
	static int pg(void* inst, const lcg_float* m, const lcg_float c, const lcg_para* p, const int n, const int k) { return static_cast<LCG_Solver*>(inst)->Progress(m, c, p, n, k); }
};
struct ClcgTramp {
	static void ax(void* inst, const lcg_complex* x, lcg_complex* y, const int n, lcg_matrix_e l, clcg_complex_e c) { static_cast<CLCG_Solver*>(inst)->AxProduct(x, y, n, l, c); }
	static int pg(void* inst, const lcg_complex* m, const lcg_float c, const clcg_para* p, const int n, const int k) { return static_cast<CLCG_Solver*>(inst)->Progress(m, c, p, n, k); }
};
struct LcgCudaTramp {
	static void ax(void* inst, cublasHandle_t cb, cusparseHandle_t cs, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int n, const int nz)
	{ static_cast<LCG_CUDA_Solver*>(inst)->AxProduct(cb, cs, x, y, n, nz); }
	static void mx(void* inst, cublasHandle_t cb, cusparseHandle_t cs, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int n, const int nz)
	{ static_cast<LCG_CUDA_Solver*>(inst)->MxProduct(cb, cs, x, y, n, nz); }   // the reference's own trampoline calls AxProduct here (solver_cuda.h:87-91)
	static int pg(void* inst, const lcg_float* m, const lcg_float c, const lcg_para* p, const int n, const int nz, const int k)
	{ return static_cast<LCG_CUDA_Solver*>(inst)->Progress(m, c, p, n, nz, k); }
};
struct ClcgCudaTramp {
	static void ax(void* inst, cublasHandle_t cb, cusparseHandle_t cs, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int n, const int nz, cusparseOperation_t op)
	{ static_cast<CLCG_CUDA_Solver*>(inst)->AxProduct(cb, cs, x, y, n, nz, op); }
	static void mx(void* inst, cublasHandle_t cb, cusparseHandle_t cs, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int n, const int nz, cusparseOperation_t op)
	{ static_cast<CLCG_CUDA_Solver*>(inst)->MxProduct(cb, cs, x, y, n, nz, op); }
	static int pg(void* inst, const cuDoubleComplex* m, const lcg_float c, const clcg_para* p, const int n, const int nz, const int k)
	{ return static_cast<CLCG_CUDA_Solver*>(inst)->Progress(m, c, p, n, nz, k); }
};
}  // namespace

void LCG_Solver::Minimize(lcg_float* m, const lcg_float* b, int x_size, lcg_solver_enum solver_id, bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, false, real_name(solver_id), [&](bool mon) {
		return lcg_solver(LcgTramp::ax, mon ? LcgTramp::pg : nullptr, m, b, x_size, &param_, this, solver_id); });
}
void LCG_Solver::MinimizePreconditioned(lcg_float* m, const lcg_float* b, int x_size, lcg_solver_enum solver_id, bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, false, real_name(LCG_PCG), [&](bool mon) {
		return lcg_solver_preconditioned(LcgTramp::ax, LcgTramp::mx, mon ? LcgTramp::pg : nullptr, m, b, x_size, &param_, this, solver_id); });
}
void LCG_Solver::MinimizeConstrained(lcg_float* m, const lcg_float* b, const lcg_float* low, const lcg_float* hig, int x_size, lcg_solver_enum solver_id,
	bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, false, real_name(solver_id), [&](bool mon) {
		return lcg_solver_constrained(LcgTramp::ax, mon ? LcgTramp::pg : nullptr, m, b, low, hig, x_size, &param_, this, solver_id); });
}

CLCG_Solver::CLCG_Solver() : param_(kDefC), inter_(1), silent_(false) {}
int CLCG_Solver::Progress(const lcg_complex*, const lcg_float converge, const clcg_para* param, const int, const int k) { return monitor(inter_, converge, param->epsilon, k); }
void CLCG_Solver::silent() { silent_ = true; }
void CLCG_Solver::set_report_interval(unsigned int inter) { inter_ = inter; }
void CLCG_Solver::set_clcg_parameter(const clcg_para& in_param) { param_ = in_param; }
void CLCG_Solver::Minimize(lcg_complex* m, const lcg_complex* b, int x_size, clcg_solver_enum solver_id, bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, true, cplx_name(solver_id), [&](bool mon) {
		return clcg_solver(ClcgTramp::ax, mon ? ClcgTramp::pg : nullptr, m, b, x_size, &param_, this, solver_id); });
}

LCG_CUDA_Solver::LCG_CUDA_Solver() : param_(kDef), inter_(1), silent_(false) {}
int LCG_CUDA_Solver::Progress(const lcg_float*, const lcg_float converge, const lcg_para* param, const int, const int, const int k)
{ return monitor(inter_, converge, param->epsilon, k); }
void LCG_CUDA_Solver::silent() { silent_ = true; }
void LCG_CUDA_Solver::set_report_interval(unsigned int inter) { inter_ = inter; }
void LCG_CUDA_Solver::set_lcg_parameter(const lcg_para& in_param) { param_ = in_param; }
void LCG_CUDA_Solver::Minimize(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, lcg_float* x, lcg_float* b, const int n_size, const int nz_size,
	lcg_solver_enum solver_id, bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, false, real_name(solver_id), [&](bool mon) {
		return lcg_solver_cuda(LcgCudaTramp::ax, mon ? LcgCudaTramp::pg : nullptr, x, b, n_size, nz_size, &param_, this, cub_handle, cus_handle, solver_id); });
}
void LCG_CUDA_Solver::MinimizePreconditioned(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, lcg_float* x, lcg_float* b, const int n_size,
	const int nz_size, lcg_solver_enum solver_id, bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, false, real_name(LCG_PCG), [&](bool mon) {
		return lcg_solver_preconditioned_cuda(LcgCudaTramp::ax, LcgCudaTramp::mx, mon ? LcgCudaTramp::pg : nullptr, x, b, n_size, nz_size, &param_, this,
			cub_handle, cus_handle, solver_id); });
}
void LCG_CUDA_Solver::MinimizeConstrained(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, lcg_float* x, const lcg_float* b, const lcg_float* low,
	const lcg_float* hig, const int n_size, const int nz_size, lcg_solver_enum solver_id, bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, false, real_name(solver_id), [&](bool mon) {
		return lcg_solver_constrained_cuda(LcgCudaTramp::ax, mon ? LcgCudaTramp::pg : nullptr, x, b, low, hig, n_size, nz_size, &param_, this, cub_handle,
			cus_handle, solver_id); });
}

CLCG_CUDA_Solver::CLCG_CUDA_Solver() : param_(kDefC), inter_(1), silent_(false) {}
int CLCG_CUDA_Solver::Progress(const cuDoubleComplex*, const lcg_float converge, const clcg_para* param, const int, const int, const int k)
{ return monitor(inter_, converge, param->epsilon, k); }
void CLCG_CUDA_Solver::silent() { silent_ = true; }
void CLCG_CUDA_Solver::set_report_interval(unsigned int inter) { inter_ = inter; }
void CLCG_CUDA_Solver::set_clcg_parameter(const clcg_para& in_param) { param_ = in_param; }
void CLCG_CUDA_Solver::Minimize(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cuDoubleComplex* x, cuDoubleComplex* b, const int n_size,
	const int nz_size, clcg_solver_enum solver_id, bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, true, cplx_name(solver_id), [&](bool mon) {
		return clcg_solver_cuda(ClcgCudaTramp::ax, mon ? ClcgCudaTramp::pg : nullptr, x, b, n_size, nz_size, &param_, this, cub_handle, cus_handle, solver_id); });
}
void CLCG_CUDA_Solver::MinimizePreconditioned(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cuDoubleComplex* x, cuDoubleComplex* b,
	const int n_size, const int nz_size, clcg_solver_enum solver_id, bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, true, cplx_name(CLCG_PCG), [&](bool mon) {
		return clcg_solver_preconditioned_cuda(ClcgCudaTramp::ax, ClcgCudaTramp::mx, mon ? ClcgCudaTramp::pg : nullptr, x, b, n_size, nz_size, &param_, this,
			cub_handle, cus_handle, solver_id); });
}

// solver_cuda.h:213-374 / solver_cuda.cu:182-298
namespace {
struct ClcgCudafTramp {
	static void ax(void* inst, cublasHandle_t cb, cusparseHandle_t cs, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int n, const int nz, cusparseOperation_t op)
	{ static_cast<CLCG_CUDAF_Solver*>(inst)->AxProduct(cb, cs, x, y, n, nz, op); }
	static void mx(void* inst, cublasHandle_t cb, cusparseHandle_t cs, cusparseDnVecDescr_t x, cusparseDnVecDescr_t y, const int n, const int nz, cusparseOperation_t op)
	{ static_cast<CLCG_CUDAF_Solver*>(inst)->MxProduct(cb, cs, x, y, n, nz, op); }
	static int pg(void* inst, const cuComplex* m, const float c, const clcg_para* p, const int n, const int nz, const int k)
	{ return static_cast<CLCG_CUDAF_Solver*>(inst)->Progress(m, c, p, n, nz, k); }
};
}  // namespace
CLCG_CUDAF_Solver::CLCG_CUDAF_Solver() : param_(kDefC), inter_(1), silent_(false) {}
int CLCG_CUDAF_Solver::Progress(const cuComplex*, const float converge, const clcg_para* param, const int, const int, const int k)
{ return monitor(inter_, converge, param->epsilon, k); }
void CLCG_CUDAF_Solver::silent() { silent_ = true; }
void CLCG_CUDAF_Solver::set_report_interval(unsigned int inter) { inter_ = inter; }
void CLCG_CUDAF_Solver::set_clcg_parameter(const clcg_para& in_param) { param_ = in_param; }
void CLCG_CUDAF_Solver::Minimize(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cuComplex* x, cuComplex* b, const int n_size,
	const int nz_size, clcg_solver_enum solver_id, bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, true, cplx_name(solver_id), [&](bool mon) {
		return clcg_solver_cuda(ClcgCudafTramp::ax, mon ? ClcgCudafTramp::pg : nullptr, x, b, n_size, nz_size, &param_, this, cub_handle, cus_handle, solver_id); });
}
void CLCG_CUDAF_Solver::MinimizePreconditioned(cublasHandle_t cub_handle, cusparseHandle_t cus_handle, cuComplex* x, cuComplex* b,
	const int n_size, const int nz_size, clcg_solver_enum solver_id, bool verbose, bool er_throw)
{
	minimize(silent_, verbose, er_throw, true, cplx_name(CLCG_PCG), [&](bool mon) {
		return clcg_solver_preconditioned_cuda(ClcgCudafTramp::ax, ClcgCudafTramp::mx, mon ? ClcgCudafTramp::pg : nullptr, x, b, n_size, nz_size, &param_, this,
			cub_handle, cus_handle, solver_id); });
}
