// engine.cu — non-template parts of the solve engine (state upload/readback, host loop, error plumbing).
#include "engine.cuh"
#include <cstring>
#include <string>
#include <mutex>
#include <cstdlib>
#include <vector>
#include <utility>

namespace lcgb200 {

static thread_local std::string g_err;
void set_error(const char* what, cudaError_t e, const char* file, int line)
{
	char buf[1024];
	snprintf(buf, sizeof(buf), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
	g_err = buf;
}
void set_error_msg(const char* msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }

Settings& settings() { static Settings s; return s; }

// PDL pays where launch gaps are a visible share of an iteration (the same regime as the CUDA graphs: up to ~8 M local
// rows); on the largest systems (7-point 512^3 on one GPU) letting the next grid become resident early costs more than the
// gap it hides (measured: BiCGSTAB 151 -> 131 it/s), so the automatic setting is decided per solve from the local size.
static thread_local bool t_pdl_active = false;
bool pdl_enabled() { return t_pdl_active; }
static void pdl_decide(size_t n_local)
{
	int mode = settings().pdl;
	if (mode < 0) { static const int env = [] { const char* e = getenv("LCGB200_PDL"); return e ? atoi(e) : -1; }(); mode = env; }
	t_pdl_active = mode > 0 || (mode < 0 && n_local <= ((size_t)8 << 20));
}

bool reference_order()
{
	int mode = settings().reference_order;
	if (mode < 0) { static const int env = [] { const char* e = getenv("LCGB200_REFERENCE_ORDER"); return e ? atoi(e) : 0; }(); mode = env; }
	return mode > 0;
}

long long spin_timeout_ms()
{
	if (settings().spin_timeout_ms >= 0) return settings().spin_timeout_ms;
	static const long long env = [] { const char* e = getenv("LCGB200_SPIN_TIMEOUT_MS"); return e ? atoll(e) : 30000ll; }();
	return env < 0 ? 30000ll : env;
}

int spmv_grid_limit(int ctas_per_sm)
{
	static int cached[64] = {0};
	int dev = 0, sms = 148;
	if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64)
	{
		if (!cached[dev])
		{
			int v = 0;
			cached[dev] = (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) ? v : 148;
		}
		sms = cached[dev];
	}
	int g = sms * ctas_per_sm;
	return g > kMaxBlocks ? kMaxBlocks : g;
}

Engine::Engine(cudaStream_t s, CsrHandle* h) : stream(s), cache(h)
{
	if (h)
	{
		comm = h->comm;
		if (!h->d_state)
		{
			LCG_CUDA_CHECK(cudaMalloc(&h->d_state, sizeof(DevState)));
			LCG_CUDA_CHECK(cudaMallocHost(&h->h_state, sizeof(DevState)));
			LCG_CUDA_CHECK(cudaMallocHost(&h->h_state2, sizeof(DevState)));
			LCG_CUDA_CHECK(cudaMalloc(&h->d_partials, sizeof(double) * kMaxBlocks * kMaxRed));
			for (int i = 0; i < 4; i++) LCG_CUDA_CHECK(cudaEventCreate(&h->ev[i]));
		}
		d_st = h->d_state; h_st = h->h_state; h_st2 = h->h_state2; d_partials = h->d_partials;
		for (int i = 0; i < 4; i++) ev[i] = h->ev[i];
		ws = (char*)h->ws; ws_cap = h->ws_bytes;
	}
	else
	{
		own_state = true;
		LCG_CUDA_CHECK(cudaMalloc(&d_st, sizeof(DevState)));
		LCG_CUDA_CHECK(cudaMallocHost(&h_st, sizeof(DevState)));
		LCG_CUDA_CHECK(cudaMallocHost(&h_st2, sizeof(DevState)));
		LCG_CUDA_CHECK(cudaMalloc(&d_partials, sizeof(double) * kMaxBlocks * kMaxRed));
		for (int i = 0; i < 4; i++) LCG_CUDA_CHECK(cudaEventCreate(&ev[i]));
	}
}

int coop_grid_limit(const void* kernel, int block)
{
	int dev = 0, sms = 148, per_sm = 1;
	cudaGetDevice(&dev);
	if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
	if (per_sm > 2) per_sm = 2;   // grid barriers get slower with more blocks; two per SM are plenty for an L2-resident system
	return sms * per_sm;
}

int coop_grid_full(const void* kernel, int block)
{
	struct Key { const void* k; int dev; };
	static thread_local std::vector<std::pair<Key, int>> cache;
	int dev = 0, sms = 148, per_sm = 1;
	cudaGetDevice(&dev);
	for (const auto& e : cache) if (e.first.k == kernel && e.first.dev == dev) return e.second;
	if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
	cache.push_back({Key{kernel, dev}, sms * per_sm});
	return sms * per_sm;
}

cudaError_t launch_coop_pdl(const void* kern, int grid, int block, void** args, cudaStream_t s)
{
	static int combo = [] { const char* e = getenv("LCGB200_VEC2_PDL"); return e ? atoi(e) : 1; }();   // 0 once refused
	if (combo && pdl_enabled())
	{
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = s;
		cudaLaunchAttribute attr[2];
		attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
		attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr; cfg.numAttrs = 2;
		const cudaError_t e = cudaLaunchKernelExC(&cfg, kern, args);
		if (e == cudaSuccess) return e;
		(void)cudaGetLastError();
		combo = 0;
		static const bool dbg = getenv("LCGB200_DEBUG_L2") != nullptr;
		if (dbg) fprintf(stderr, "[lcgb200] cooperative + programmatic launch refused (%s): plain cooperative launches from here on\n", cudaGetErrorString(e));
	}
	return cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3((unsigned)block), args, 0, s);
}

bool fuse_vec2(size_t n_local)
{
	int mode = settings().fuse_vec2;
	if (mode < 0) { static const int env = [] { const char* e = getenv("LCGB200_FUSE_VEC2"); return e ? atoi(e) : -1; }(); mode = env; }
	if (mode == 0) return false;
	if (mode > 0) return true;
	return n_local <= ((size_t)8 << 20);
}

cudaEvent_t Engine::prof_begin(int cls)
{
	if (timed_used == timed.size())
	{
		if (timed.size() >= 8192) return nullptr;
		Timed t; t.cls = cls;
		if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) return nullptr;
		timed.push_back(t);
	}
	Timed& t = timed[timed_used++];
	t.cls = cls;
	cudaEventRecord(t.a, stream);
	return t.b;
}

void Engine::prof_collect(double* ms, int* count)
{
	// Launches enqueued after the device set `done` return at their first instruction (~3 us): they are not part of
	// the algorithm, so only launches lasting at least a quarter of their class's longest one are counted.
	ms[0] = ms[1] = 0.0; count[0] = count[1] = 0;
	std::vector<float> dur(timed_used, 0.f);
	float mx[2] = {0.f, 0.f};
	for (size_t i = 0; i < timed_used; i++)
	{
		if (cudaEventElapsedTime(&dur[i], timed[i].a, timed[i].b) != cudaSuccess) dur[i] = 0.f;
		if (dur[i] > mx[timed[i].cls]) mx[timed[i].cls] = dur[i];
	}
	for (size_t i = 0; i < timed_used; i++)
	{
		const int c = timed[i].cls;
		if (dur[i] >= 0.25f * mx[c] && dur[i] > 0.f) { ms[c] += dur[i]; count[c]++; }
	}
}

Engine::~Engine()
{
	l2_window_end();
	for (auto& t : timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
	if (own_state)
	{
		cudaFree(d_st); cudaFreeHost(h_st); cudaFreeHost(h_st2); cudaFree(d_partials);
		for (int i = 0; i < 4; i++) if (ev[i]) cudaEventDestroy(ev[i]);
	}
	if (own_ws && ws) cudaFree(ws);
}

// Mid-size systems (10^6-10^7 rows per GPU — BASELINE configs[2], the per-GPU slabs of configs[3] on 8 GPUs): the work
// vectors would fit the 126 MB L2 but the matrix stream (hundreds of MB per SpMV, already tagged evict-first) keeps washing
// them out.  For the duration of a solve the tail of the arena — as many WHOLE work vectors as fit the device's persisting
// carve-out (79 MiB on B200), counted from the last one allocated: the SpMV's input and output come last in every solver —
// becomes a persisting access-policy window of the solve's stream (CUDA graphs capture it into their kernel nodes).
// Measured on one B200: 27-point 128^3 PCG 6784 -> 7179 it/s, 7-point 128^3 CG 16434 -> 16844 it/s.  A window LARGER than
// the carve-out (hit ratio < 1) is harmful (27-point 160^3: -10 %, 256^3: -30 %), hence whole vectors only.  The carve-out is
// a device-wide setting: it is released (cudaCtxResetPersistingL2Cache) when the solve ends.
// lcgb200_set_l2_persist / LCGB200_L2_PERSIST: -1 automatic (at least two vectors and about half of them fit), 0 off,
// 1 whenever at least one vector fits.
void Engine::l2_window_begin()
{
	int mode = settings().l2_persist;
	if (mode < 0) { static const int env = [] { const char* e = getenv("LCGB200_L2_PERSIST"); return e ? atoi(e) : -1; }(); mode = env; }
	if (mode == 0 || !ws || !l2_unit || ws_need < l2_from + l2_unit || stream == nullptr || stream == cudaStreamLegacy) return;
	if (mode < 0 && (!cache || n_local <= (size_t)kSmallRows)) return;   // automatic: built-in operator, not the cache-resident regime
	int dev = 0, max_persist = 0, max_window = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) return;
	cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
	cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
	if (max_persist <= 0 || max_window <= 0) return;
	const size_t n_vecs = (ws_need - l2_from) / l2_unit;
	size_t fit = (size_t)(max_persist < max_window ? max_persist : max_window) / l2_unit;
	if (fit > n_vecs) fit = n_vecs;
	if (fit == 0 || (mode < 0 && (fit < 2 || 2 * fit < n_vecs - 1))) return;   // measured: 2 of 4 vectors +3.5 %, 1 of 4 -1 %
	const size_t bytes = fit * l2_unit;
	// the carve-out is taken from the L2 every other access uses: remember the previous setting and put it back when the solve
	// ends (left in place it costs later solves of LARGE systems a quarter of their SpMV bandwidth: 7-point 512^3 on 8 GPUs
	// 0.293 -> 0.378 ms per SpMV, measured)
	if (cudaDeviceGetLimit(&l2_prev_limit, cudaLimitPersistingL2CacheSize) != cudaSuccess) { (void)cudaGetLastError(); l2_prev_limit = 0; }
	if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) != cudaSuccess) { (void)cudaGetLastError(); return; }
	cudaStreamAttrValue v = {};
	v.accessPolicyWindow.base_ptr = ws + (ws_need - bytes);
	v.accessPolicyWindow.num_bytes = bytes;
	v.accessPolicyWindow.hitRatio = 1.0f;
	v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
	v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
	if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) { (void)cudaGetLastError(); return; }
	l2_window_set = true;
	static const bool dbg = getenv("LCGB200_DEBUG_L2") != nullptr;
	if (dbg) fprintf(stderr, "[lcgb200] L2 window: %zu of %zu work vectors (%zu B), max persisting %d B, max window %d B\n", fit, n_vecs, bytes, max_persist, max_window);
}

void Engine::l2_window_end()
{
	if (!l2_window_set) return;
	cudaStreamAttrValue v = {};
	v.accessPolicyWindow.num_bytes = 0;
	cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v);
	cudaCtxResetPersistingL2Cache();
	if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, l2_prev_limit) != cudaSuccess) (void)cudaGetLastError();
	l2_window_set = false;
}

void Engine::reserve(size_t bytes)
{
	bytes = (bytes + 255) & ~size_t(255);
	ws_off = 0;
	ws_need = bytes;
	if (cache)
	{
		if (cache->ws_bytes < bytes)
		{
			if (cache->ws) { LCG_CUDA_CHECK(cudaFree(cache->ws)); cache->ws = nullptr; cache->ws_bytes = 0; }
			LCG_CUDA_CHECK(cudaMalloc(&cache->ws, bytes));
			cache->ws_bytes = bytes;
		}
		ws = (char*)cache->ws; ws_cap = cache->ws_bytes;
	}
	else
	{
		if (ws && own_ws) { cudaFree(ws); ws = nullptr; }
		LCG_CUDA_CHECK(cudaMalloc((void**)&ws, bytes));
		ws_cap = bytes; own_ws = true;
	}
}

void Engine::start(const DevState& init)
{
	*h_st = init;
	h_st->multi = multi() ? (p2p() ? 2 : 1) : 0;
	h_st->comm = p2p() ? cache->p2p_dev() : nullptr;
	// a progress callback or a host-side operator puts the host into every iteration: a rank may then legitimately be
	// seconds late (slow callback, debugger), so the peers wait 20x longer before they give up
	h_st->spin_timeout_ns = (unsigned long long)spin_timeout_ms() * 1000000ull * ((pf || sync_each) ? 20ull : 1ull);
	pushed_vec = nullptr;
	pdl_decide(n_local);
	l2_window_begin();
	LCG_CUDA_CHECK(cudaMemcpyAsync(d_st, h_st, sizeof(DevState), cudaMemcpyHostToDevice, stream));
	LCG_CUDA_CHECK(cudaEventRecord(ev[2], stream));
	seen_checks = 0;
	launches = 0; spmv_launches = 0;
	profiling = settings().profile != 0; timed_used = 0;
	final_ret = RC_UNKNOWN;
}

void Engine::read_state()
{
	LCG_CUDA_CHECK(cudaMemcpyAsync(h_st, d_st, sizeof(DevState), cudaMemcpyDeviceToHost, stream));
	LCG_CUDA_CHECK(cudaStreamSynchronize(stream));
}

double Engine::device_ms()
{
	LCG_CUDA_CHECK(cudaEventRecord(ev[3], stream));
	LCG_CUDA_CHECK(cudaEventSynchronize(ev[3]));
	float ms = 0.f;
	LCG_CUDA_CHECK(cudaEventElapsedTime(&ms, ev[2], ev[3]));
	return (double)ms;
}

bool Engine::sync_always()
{
	read_state();
	if (h_st->checks != seen_checks)
	{
		seen_checks = h_st->checks;
		static const bool dbg = getenv("LCGB200_DEBUG_SCALARS") != nullptr;
		if (dbg)
			fprintf(stderr, "[lcgb200] k=%d res=%.6e rho=(%.6e,%.6e) alpha=(%.6e,%.6e) beta=(%.6e,%.6e) omega=(%.6e,%.6e) mm=%.6e rr=%.6e\n",
				h_st->k_report, h_st->residual, h_st->sc[SC_RHO], h_st->sc[SC_RHO_I], h_st->sc[SC_ALPHA], h_st->sc[SC_ALPHA_I],
				h_st->sc[SC_BETA], h_st->sc[SC_BETA_I], h_st->sc[SC_OMEGA], h_st->sc[SC_OMEGA_I], h_st->sc[SC_MMOD], h_st->sc[SC_RMOD]);
		if (pf)
		{
			int stop = pf(h_st->residual, h_st->k_report);
			// the "already optimised" call ignores the callback's return value (lcg.cpp:189-192)
			if (stop && !(h_st->done && h_st->ret == RC_ALREADY)) { final_ret = RC_STOP; return true; }
		}
	}
	if (h_st->done) { final_ret = h_st->ret; return true; }
	return false;
}

bool Engine::sync_point()
{
	if (!pf) return false;
	return sync_always();
}

namespace {
struct GraphGuard {
	cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
	void reset() { if (exec) cudaGraphExecDestroy(exec); if (graph) cudaGraphDestroy(graph); exec = nullptr; graph = nullptr; }
	~GraphGuard() { reset(); }
};
// graphs pay when launch gaps are a visible share of an iteration: up to ~8 M local rows (an iteration of a few hundred
// microseconds); lcgb200_set_graphs / LCGB200_GRAPHS = 0 | 1 overrides
bool use_graphs(size_t n_local)
{
	int mode = settings().graphs;
	if (mode < 0) { static const int env = [] { const char* e = getenv("LCGB200_GRAPHS"); return e ? atoi(e) : -1; }(); mode = env; }
	if (mode == 0) return false;
	if (mode > 0) return true;
	return n_local <= (size_t)8 << 20;
}
}  // namespace

int Engine::run(const std::function<bool()>& iterate, const std::function<void(int)>& batch)
{
	if (pf || sync_each)
	{	// one host round trip per loop head, exactly like the reference
		while (true)
		{
			if (sync_always()) break;
			if (batch) batch(1);
			else if (iterate()) break;
			LCG_CUDA_CHECK(cudaPeekAtLastError());
		}
		LCG_CUDA_CHECK(cudaStreamSynchronize(stream));
		return final_ret;
	}
	// No callback: enqueue `poll` iterations per batch, read the state back asynchronously, and look at the
	// PREVIOUS batch's flag while the current one runs.  Kernels launched after `done` return immediately.
	// fused launches cost ~one launch per batch: poll less often (an iteration of a cache-resident system takes microseconds)
	// (graph-sized systems: an iteration lasts tens of microseconds, so four times as many per batch — the gap between two graph
	// launches is then a percent of the batch, and iterations enqueued past convergence are no-ops)
	const bool want_graph_pre = !batch && capturable && !profiling && (!multi() || p2p()) && stream != nullptr && stream != cudaStreamLegacy && use_graphs(n_local);
	const int poll = (settings().poll > 0 ? settings().poll : 1) * (batch ? 8 : (want_graph_pre ? 4 : 1));
	DevState* slot[2] = {h_st, h_st2};
	bool pending[2] = {false, false};
	int b = 0;
	// the state after the init kernels (already optimised?)
	LCG_CUDA_CHECK(cudaMemcpyAsync(slot[b], d_st, sizeof(DevState), cudaMemcpyDeviceToHost, stream));
	LCG_CUDA_CHECK(cudaEventRecord(ev[b], stream));
	pending[b] = true;
	b ^= 1;
	// CUDA graph of one batch: every kernel argument of an iteration is a constant of the solve (the scalars live in
	// DevState), so the second batch is captured once and replayed — kernel-to-kernel gaps shrink to the graph's
	// pre-resolved dependencies, which matters when an iteration lasts tens of microseconds (mid-size systems, the
	// per-GPU slabs of a partitioned solve).  The first batch runs uncaptured so that one-time launch set-up is done.
	GraphGuard gg;
	// (the legacy default stream cannot be captured: the API layer moves built-in-operator solves to a stream of the handle)
	const bool want_graph = want_graph_pre;
	int batches = 0, per_batch_launches = 0, per_batch_spmv = 0;
	while (true)
	{
		bool ended = false;
		if (batch) batch(poll);
		else if (gg.exec)
		{
			LCG_CUDA_CHECK(cudaGraphLaunch(gg.exec, stream));
			launches += per_batch_launches; spmv_launches += per_batch_spmv;
		}
		else if (want_graph && batches == 1)
		{
			const int l0 = launches, s0 = spmv_launches;
			cudaError_t ce = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
			if (ce == cudaSuccess)
			{
				for (int i = 0; i < poll; i++) iterate();
				ce = cudaStreamEndCapture(stream, &gg.graph);
			}
			if (ce == cudaSuccess && gg.graph) ce = cudaGraphInstantiate(&gg.exec, gg.graph, 0);
			if (ce != cudaSuccess || !gg.exec)
			{	// capture refused (e.g. a user stream in a state that forbids it): forget it and carry on with plain launches
				(void)cudaGetLastError();
				gg.reset();
				launches = l0; spmv_launches = s0;
				for (int i = 0; i < poll && !ended; i++) ended = iterate();
				batches = 2;   // do not try again
			}
			else
			{
				per_batch_launches = launches - l0; per_batch_spmv = spmv_launches - s0;
				LCG_CUDA_CHECK(cudaGraphLaunch(gg.exec, stream));
			}
		}
		else for (int i = 0; i < poll && !ended; i++) ended = iterate();
		batches++;
		if (ended) break;   // a host-driven step (SPG) saw the end itself
		LCG_CUDA_CHECK(cudaPeekAtLastError());   // a launch that failed would never set `done`
		LCG_CUDA_CHECK(cudaMemcpyAsync(slot[b], d_st, sizeof(DevState), cudaMemcpyDeviceToHost, stream));
		LCG_CUDA_CHECK(cudaEventRecord(ev[b], stream));
		pending[b] = true;
		const int o = b ^ 1;
		if (pending[o])
		{
			LCG_CUDA_CHECK(cudaEventSynchronize(ev[o]));
			pending[o] = false;
			if (slot[o]->done) { final_ret = slot[o]->ret; break; }
		}
		b = o;
	}
	LCG_CUDA_CHECK(cudaStreamSynchronize(stream));
	if (final_ret == RC_UNKNOWN || !pf)
	{	// take the authoritative final state
		LCG_CUDA_CHECK(cudaMemcpy(h_st, d_st, sizeof(DevState), cudaMemcpyDeviceToHost));
		if (h_st->done) final_ret = h_st->ret;
	}
	return final_ret;
}

}  // namespace lcgb200
