// Host side of the row-pattern operator copy (csr.cuh "row-pattern operator"): choice of the matrix-wide stride S and the
// grouping of a pattern's entries into chains of offsets S apart.  Plain C++ (no CUDA) so that tests/cxx/pat_chain_check.cpp
// can run the decomposition and an emulation of the kernel's tiling on the CPU.
#pragma once
#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>

namespace lcgb200 {

constexpr int kPatChainLenH = 3;
struct PatChainH { double v[kPatChainLenH]; int off; int m; };
static_assert(sizeof(PatChainH) == 32, "layout of csr.cuh: PatChain");

// entries = one row as (col - row, value) pairs in row order.  Greedy: walk the entries by ascending offset, every
// entry not yet used starts a chain and pulls in the first unused entries at +S and +2S.  The chain holding offset 0
// (the diagonal) is moved to the front; *t0 = its index inside that chain, -1 when the row has no diagonal entry.
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
			if ((long long)chains[c].off + (long long)t * S == 0) { std::swap(chains[0], chains[c]); *t0 = t; break; }
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
					for (int t = 0; t < chains[s][c].m; t++)
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

}  // namespace lcgb200
