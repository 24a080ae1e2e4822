// Host side of the row-pattern operator copy (csr.cuh "row-pattern operator"): choice of the matrix-wide stride S, the
// grouping of a pattern's entries into chains of offsets S apart, the subset masks, and the plan of the plane-marching
// kernel.  Plain C++ (no CUDA) so that tests/cxx/pat_chain_check.cpp can run the decomposition and an emulation of the
// kernels' tiling on the CPU.
#pragma once
#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>

namespace lcgb200 {

constexpr int kPatChainLenH = 3;
// m = entries (1..3) | line shift << 4 | position inside the window line << 8 | plane << 16 (the last three only for the
// chains of the marching kernel's geometry pattern; the other users look at m & 3)
struct PatChainH { double v[kPatChainLenH]; int off; int m; };
static_assert(sizeof(PatChainH) == 32, "layout of csr.cuh: PatChain");

// entries = one row as (col - row, value) pairs in row order.  Greedy: walk the entries by ascending offset, every
// entry not yet used starts a chain and pulls in the first unused entries at +S and +2S.  Chains come out ordered by
// their first offset, except that the chain holding offset 0 (the diagonal) is moved to the END: the kernels take x[row]
// for the fused dot products from the registers of the last chain they process.  *t0 = the diagonal's index inside that
// chain, -1 when the row has no diagonal entry.
inline void pat_build_chains(const std::vector<std::pair<int, double>>& entries, int S, std::vector<PatChainH>& chains, int* t0)
{
	const size_t len = entries.size();
	std::vector<size_t> order(len);
	for (size_t i = 0; i < len; i++) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return entries[a].first < entries[b].first; });
	std::vector<char> used(len, 0);
	chains.clear();
	for (size_t a = 0; a < len; a++)
	{
		if (used[a]) continue;
		used[a] = 1;
		PatChainH c{{entries[order[a]].second, 0.0, 0.0}, entries[order[a]].first, 1};
		for (size_t b = a + 1; b < len && c.m < kPatChainLenH; b++)
		{
			if (used[b]) continue;
			const long long want = (long long)c.off + (long long)c.m * S;
			if ((long long)entries[order[b]].first > want) break;
			if ((long long)entries[order[b]].first == want) { used[b] = 1; c.v[c.m++] = entries[order[b]].second; }
		}
		chains.push_back(c);
	}
	*t0 = -1;
	for (size_t c = 0; c < chains.size() && *t0 < 0; c++)
		for (int t = 0; t < chains[c].m; t++)
			if ((long long)chains[c].off + (long long)t * S == 0)
			{
				const PatChainH d = chains[c];
				chains.erase(chains.begin() + (ptrdiff_t)c);
				chains.push_back(d);
				*t0 = t;
				break;
			}
}

// S = the difference between two offsets of the pattern (>= 32 so that a warp's 32 consecutive rows stay inside one
// super-row of S rows, and R*S <= n_rows so that full work items exist) that leaves the fewest chains; ties go to the
// stride that is itself an offset of the pattern (the line stride nx of a grid: rows S apart then sit at the same place of
// their grid lines and share their pattern; a difference such as nx - 1 would walk them into the boundary), then to the
// smaller stride (the R rows of a thread stay close in memory).  fallback when no difference qualifies or none saves a load.
inline int pat_pick_stride(const std::vector<std::pair<int, double>>& entries, long long n_rows, int R, int fallback)
{
	std::vector<int> offs;
	for (const auto& e : entries) offs.push_back(e.first);
	std::sort(offs.begin(), offs.end());
	offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
	std::vector<int> cand;
	for (size_t i = 0; i < offs.size(); i++)
		for (size_t j = i + 1; j < offs.size(); j++)
		{
			const long long d = (long long)offs[j] - (long long)offs[i];
			if (d >= 32 && d * R <= n_rows) cand.push_back((int)d);
		}
	std::sort(cand.begin(), cand.end());
	cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
	int best = fallback; size_t best_chains = entries.size(); bool best_is_offset = false;
	std::vector<PatChainH> tmp; int t0;
	for (int d : cand)
	{
		pat_build_chains(entries, d, tmp, &t0);
		const bool is_offset = std::binary_search(offs.begin(), offs.end(), d);
		if (tmp.size() < best_chains || (tmp.size() == best_chains && best != fallback && is_offset && !best_is_offset))
		{
			best_chains = tmp.size(); best = d; best_is_offset = is_offset;
		}
	}
	return best;
}

// Rows whose pattern is a SUBSET of a longer one (the face / edge / corner rows of a stencil are its interior row minus the
// entries that would leave the grid) can ride that pattern's chains with the missing entries masked out, so a thread whose
// R rows do not all share one pattern — a column of grid points that starts on the boundary — still shares its loads.
// For every pattern p: sup[p] = the longest pattern s (at most kPatMaskBits / kPatChainLenH chains) that holds every entry
// of p with the same offset and bitwise the same value; mask[p] bit 3 c + t = entry t of chain c of s is present in p.
// sup[p] = p with all its own bits when nothing longer fits; mask[p] = 0 when p has too many chains for a 64-bit mask.
constexpr int kPatMaskBits = 64;
inline void pat_build_masks(const std::vector<std::vector<std::pair<int, double>>>& rows, const std::vector<std::vector<PatChainH>>& chains, int S,
	std::vector<int>& sup, std::vector<unsigned long long>& mask)
{
	const size_t np = rows.size();
	sup.assign(np, 0); mask.assign(np, 0ull);
	auto same_bits = [](double a, double b) { unsigned long long x, y; std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8); return x == y; };
	for (size_t p = 0; p < np; p++)
	{
		sup[p] = (int)p;
		size_t best_len = 0;
		for (size_t s = 0; s < np; s++)
		{
			if (chains[s].size() * kPatChainLenH > (size_t)kPatMaskBits) continue;
			if (rows[s].size() < rows[p].size() || rows[s].size() <= best_len) continue;
			unsigned long long m = 0ull; bool ok = true;
			for (const auto& e : rows[p])
			{
				bool found = false;
				for (size_t c = 0; c < chains[s].size() && !found; c++)
					for (int t = 0; t < (chains[s][c].m & 3); t++)
					{
						const unsigned long long bit = 1ull << (c * kPatChainLenH + (size_t)t);
						if (!(m & bit) && (long long)chains[s][c].off + (long long)t * S == (long long)e.first && same_bits(chains[s][c].v[t], e.second))
						{
							m |= bit; found = true; break;
						}
					}
				if (!found) { ok = false; break; }
			}
			if (ok) { best_len = rows[s].size(); sup[p] = (int)s; mask[p] = m; }
		}
	}
}

// ---- the plane-marching kernel's plan (csr.cuh: k_spmv_pat_march) -------------------------------------------------------
// The chains of ONE pattern (the geometry pattern: the interior row of a stencil, superset of the boundary rows) are laid
// onto PLANES: a plane is a window of `nlines` segments of x, S apart, each 32 wx + kPatSpan values wide, that starts at
// row + o_min[plane]; chain c reads its values at line shift_c + u, position d_c (off_c = o_min + shift_c S + d_c).
// The planes' first offsets must be an arithmetic progression with step S2 (the plane stride nx ny of a grid): a thread
// block then marches along S2 — block item k + 1 lies S2 rows behind block item k — and the window of plane p of item k
// IS the window of plane p - 1 of item k + 1: one new window per item instead of G.
constexpr int kPatSpanH = 8;          // extra values per window line: positions d in [0, 8]
constexpr int kPatMaxShiftH = 2;      // line shift of a chain inside its plane
constexpr int kPatMaxPlanesH = 4;
struct PatMarchH {
	int ok = 0;
	int G = 0;          // planes
	int S2 = 0;         // rows between consecutive planes
	int o0 = 0;         // first offset of plane 0 (even)
	int nlines = 0;     // lines per window: wy R + max(shift + m - 1)
	int wx = 0, wy = 0; // warps of a block along the row index / along S
	int dAb = 0;        // block items (of wy R S rows) between consecutive items of a march: S2 / (wy R S)
	int group_begin[kPatMaxPlanesH + 1] = {0, 0, 0, 0, 0};   // chains [group_begin[g], group_begin[g + 1]) read plane group_plane[g]
	int group_plane[kPatMaxPlanesH] = {0, 0, 0, 0};          // table order: the diagonal's plane last
};

// chains: the geometry pattern's (diagonal chain last, as pat_build_chains leaves them); reordered plane by plane on
// success (diagonal's plane last, the diagonal chain still the very last) with the placement packed into m
inline bool pat_plan_march(std::vector<PatChainH>& chains, int t0, int S, long long n_rows, long long n_cols, int R, PatMarchH& plan)
{
	plan = PatMarchH();
	if (S % 128 != 0 || (n_cols & 1) || chains.empty()) return false;
	const int wx = (S % 256 == 0) ? 8 : 4, wy = 8 / wx;
	const size_t nc = chains.size();
	std::vector<size_t> order(nc);
	for (size_t i = 0; i < nc; i++) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return chains[a].off < chains[b].off; });
	// S2 = a difference between two chain offsets that is a multiple of the rows of a block item; every chain then sits on
	// plane p = nearest multiple of S2, at r = off - p S2 inside it; all planes share the same first offset o_base (even, so
	// that every window line starts on a 16-byte boundary: S, S2 and the row bases are even)
	std::vector<long long> cand, entry_offs;
	for (size_t c = 0; c < nc; c++) for (int t = 0; t < (chains[c].m & 3); t++) entry_offs.push_back((long long)chains[c].off + (long long)t * S);
	for (size_t i = 0; i < entry_offs.size(); i++)
		for (size_t j = 0; j < entry_offs.size(); j++)
		{
			const long long d = entry_offs[j] - entry_offs[i];
			if (d > 4LL * S && d % ((long long)wy * R * S) == 0 && d <= n_rows) cand.push_back(d);
		}
	std::sort(cand.begin(), cand.end());
	cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
	std::vector<int> plane(nc, -1), shift(nc, 0), dpos(nc, 0);
	long long S2 = 0, o_base = 0; int G = 0;
	for (long long c2 : cand)
	{
		std::vector<long long> pl(nc), r(nc);
		long long rmin = 0, pmin = 0, pmax = 0;
		for (size_t c = 0; c < nc; c++)
		{
			const long long off = chains[c].off;
			pl[c] = (off >= 0 ? (off + c2 / 2) / c2 : -((-off + c2 / 2) / c2));
			r[c] = off - pl[c] * c2;
			if (c == 0) { rmin = r[c]; pmin = pmax = pl[c]; }
			rmin = std::min(rmin, r[c]); pmin = std::min(pmin, pl[c]); pmax = std::max(pmax, pl[c]);
		}
		if (pmax - pmin + 1 > kPatMaxPlanesH || pmax == pmin) continue;
		long long ob = rmin - 1;
		ob -= ((ob % 2) + 2) % 2;
		bool ok = true;
		for (size_t c = 0; c < nc && ok; c++)
		{
			const long long q = r[c] - ob, sh = q / S, d = q - sh * S;
			if (sh > kPatMaxShiftH || d > kPatSpanH) ok = false;
			plane[c] = (int)(pl[c] - pmin); shift[c] = (int)sh; dpos[c] = (int)d;
		}
		if (!ok) continue;
		S2 = c2; G = (int)(pmax - pmin + 1); o_base = pmin * c2 + ob;
		break;
	}
	if (S2 == 0) return false;
	std::vector<int> o_min((size_t)G);
	for (int p = 0; p < G; p++) o_min[(size_t)p] = (int)(o_base + (long long)p * S2);
	int extra_lines = 0;
	for (size_t c = 0; c < nc; c++) extra_lines = std::max(extra_lines, shift[c] + (chains[c].m & 3) - 1);
	// table order: plane by plane, the diagonal's plane (the plane of the last chain, if it holds the diagonal) last
	const int diag_plane = t0 >= 0 ? plane[nc - 1] : -1;
	std::vector<PatChainH> out;
	int g = 0;
	for (int pass = 0; pass < 2; pass++)
		for (int p = 0; p < G; p++)
		{
			if ((pass == 0) == (p == diag_plane)) continue;
			plan.group_begin[g] = (int)out.size(); plan.group_plane[g] = p;
			for (size_t a = 0; a < nc; a++)
			{
				const size_t c = order[a];
				if (plane[c] != p || (t0 >= 0 && c == nc - 1)) continue;
				PatChainH ch = chains[c];
				ch.m = (ch.m & 3) | (shift[c] << 4) | (dpos[c] << 8) | (p << 16);
				out.push_back(ch);
			}
			if (p == diag_plane)
			{
				PatChainH ch = chains[nc - 1];
				ch.m = (ch.m & 3) | (shift[nc - 1] << 4) | (dpos[nc - 1] << 8) | (p << 16);
				out.push_back(ch);
			}
			g++;
		}
	plan.group_begin[G] = (int)out.size();
	chains.swap(out);
	plan.ok = 1; plan.G = G; plan.S2 = (int)S2; plan.o0 = o_min[0]; plan.nlines = wy * R + extra_lines; plan.wx = wx; plan.wy = wy;
	plan.dAb = (int)(S2 / ((long long)wy * R * S));
	return true;
}

// The marches: block items (Ab, ibb) = rows [Ab wy R S, (Ab + 1) wy R S) x positions [32 wx ibb, 32 wx (ibb + 1)) of every
// super-row; a column = the items Ab0 + k dAb of one ibb.  Columns are cut into equal segments (a segment pays G - 1 extra
// windows) so that there are about `target` segments in all — a multiple of the resident blocks, each of which then walks
// the same number of items — but never shorter than min_len items.  seg = {first Ab, ibb, items, 0}
struct PatSegH { int ab, ibb, len, pad; };
inline void pat_build_segments(const PatMarchH& plan, long long n_rows, int S, int R, int target, int min_len, std::vector<PatSegH>& segs)
{
	segs.clear();
	const long long n_super = (n_rows + S - 1) / S, n_a = (n_super + R - 1) / R;
	const long long n_ab = (n_a + plan.wy - 1) / plan.wy;
	const int nibb = S / (32 * plan.wx);
	const long long n_columns = std::min<long long>(plan.dAb, n_ab) * nibb;
	for (long long ab0 = 0; ab0 < std::min<long long>(plan.dAb, n_ab); ab0++)
	{
		const long long col_len = (n_ab - ab0 + plan.dAb - 1) / plan.dAb;
		long long n_seg = std::max<long long>(1, target / std::max<long long>(1, n_columns));
		n_seg = std::max<long long>(1, std::min<long long>(n_seg, col_len / std::max(1, min_len)));
		for (int ibb = 0; ibb < nibb; ibb++)
			for (long long s = 0; s < n_seg; s++)
			{
				const long long k0 = col_len * s / n_seg, k1 = col_len * (s + 1) / n_seg;
				if (k1 > k0) segs.push_back(PatSegH{(int)(ab0 + k0 * plan.dAb), ibb, (int)(k1 - k0), 0});
			}
	}
}


// ---- the box kernel's plan (csr.cuh: k_spmv_pat_box) --------------------------------------------------------------------
// Geometry pattern = a DENSE BOX: G groups (planes) of three chains each — offsets e_g - 1, e_g, e_g + 1 around an EVEN
// centre e_g, three entries S apart per chain (a 27-point stencil: 3 planes x (dx = -1, 0, +1) x (dy = -1, 0, +1)).  A
// thread then owns TWO neighbouring rows per line (an aligned pair) x R lines, reads each window line as one 16-byte pair
// + the two values left and right of it, and serves 18 products per line and group out of 4 loaded values.  The rows on
// the boundary of the grid are the box minus whole slices: no left / right column (dx), no lower / upper line (dy), no
// group (dz); `rowflags[p]` says which slices pattern p lacks (0x80 = p is no such sub-box of the geometry pattern).
constexpr int kPatBoxMaxGroups = 3;   // flag bits 4..6: the thread flags stay below 0x80
constexpr int kBoxHDropL = 1, kBoxHDropR = 2, kBoxHDropLow = 4, kBoxHDropHigh = 8, kBoxHDropG0 = 16;   // | kBoxHDropG0 << g (csr.cuh: kBoxDrop*)
struct PatBoxH {
	int ok = 0, G = 0;
	int center[kPatBoxMaxGroups] = {0, 0, 0};
	double coef[kPatBoxMaxGroups][3][3] = {};   // [group][dx + 1][line j]: the value at offset center + (dx) + j S
	int chain_of[kPatBoxMaxGroups][3] = {};     // chain index (in the geometry pattern's table) of (group, dx + 1)
};

// chains = the geometry pattern's (any order).  Fills plan and, with the subset masks of all patterns over the geometry
// pattern's chains (pat_build_masks), the row flags of every pattern.
inline bool pat_plan_box(const std::vector<PatChainH>& chains, int S, int gpat, const std::vector<int>& sup, const std::vector<unsigned long long>& mask,
	PatBoxH& plan, std::vector<unsigned char>& rowflags)
{
	plan = PatBoxH();
	const size_t nc = chains.size();
	if ((S & 1) || nc == 0 || nc % 3 != 0 || nc / 3 > (size_t)kPatBoxMaxGroups) return false;
	std::vector<int> used(nc, 0);
	int G = 0;
	for (size_t c = 0; c < nc; c++)
	{	// group centres: the chains with an even first offset
		if ((chains[c].m & 3) != 3) return false;
		if (chains[c].off & 1) continue;
		int left = -1, right = -1;
		for (size_t d = 0; d < nc; d++)
		{
			if (chains[d].off == chains[c].off - 1) left = (int)d;
			if (chains[d].off == chains[c].off + 1) right = (int)d;
		}
		if (left < 0 || right < 0 || used[(size_t)left] || used[(size_t)right] || used[c] || G >= kPatBoxMaxGroups) return false;
		used[c] = used[(size_t)left] = used[(size_t)right] = 1;
		plan.center[G] = chains[c].off;
		const int idx[3] = {left, (int)c, right};
		for (int dx = 0; dx < 3; dx++)
		{
			plan.chain_of[G][dx] = idx[dx];
			for (int j = 0; j < 3; j++) plan.coef[G][dx][j] = chains[(size_t)idx[dx]].v[j];
		}
		G++;
	}
	for (size_t c = 0; c < nc; c++) if (!used[c]) return false;
	plan.G = G;
	// row flags: pattern p = the box minus the slices of one flag combination?
	unsigned long long full = 0ull;
	for (size_t c = 0; c < nc; c++) full |= 7ull << (3 * c);
	auto combo_mask = [&](int f) {
		unsigned long long m = full;
		for (int g = 0; g < G; g++)
			for (int dx = 0; dx < 3; dx++)
			{
				const int c = plan.chain_of[g][dx];
				unsigned long long drop = 0ull;
				if ((f & (kBoxHDropG0 << g)) || (dx == 0 && (f & kBoxHDropL)) || (dx == 2 && (f & kBoxHDropR))) drop = 7ull;
				if (f & kBoxHDropLow) drop |= 1ull;
				if (f & kBoxHDropHigh) drop |= 4ull;
				m &= ~(drop << (3 * c));
			}
		return m;
	};
	rowflags.assign(sup.size(), 0x80);
	const int n_combo = 16 << G;
	for (size_t p = 0; p < sup.size(); p++)
	{
		if (sup[p] != gpat || mask[p] == 0ull) continue;
		for (int f = 0; f < n_combo && f < 0x80; f++)
			if (combo_mask(f) == mask[p]) { rowflags[p] = (unsigned char)f; break; }
	}
	plan.ok = 1;
	return true;
}

}  // namespace lcgb200
