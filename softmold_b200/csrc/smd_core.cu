// libsoftmold_b200.so -- C ABI implementation (include/softmold_b200.h) on top of the sm_100a kernels.
// Host-side orchestration only; there is no CPU compute path: without a usable CUDA device every compute entry
// point returns SMD_ERR_CUDA.
#include <cmath>
#include <climits>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <nvtx3/nvToolsExt.h>

#include "smd_kernels.cuh"

using namespace smd;

static std::string g_create_error;

#define CK(call)                                                                                          \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) {                                                                          \
			ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
			return SMD_ERR_CUDA;                                                                          \
		}                                                                                                 \
	} while (0)

#define REQUIRE(cond, msg)                                                                                \
	do {                                                                                                  \
		if (!(cond)) { ctx->err = (msg); return SMD_ERR_ARG; }                                            \
	} while (0)

static inline int nblk(long long n, int tpb) { return (int)std::max<long long>(1, (n + tpb - 1) / tpb); }

#define LAUNCH(kernel, grid, block, smem, ...)                                                            \
	do {                                                                                                  \
		kernel<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__);                                    \
		ctx->launches++;                                                                                  \
	} while (0)

// LAUNCHP: like LAUNCH, but with the programmatic-stream-serialization attribute when the context allows it (SMD_PDL, no
// profiling events in the stream): the kernel may become resident before its predecessor has completed and MUST order
// itself on the device (pdl_prologue(), or k_chain_kick's completion words).
template <class... KArgs, class... Args>
static cudaError_t launch_ex(bool pdl, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define LAUNCHP(kernel, grid, block, smem, ...)                                                           \
	do {                                                                                                  \
		CK(launch_ex(ctx->pdl_chain && ctx->prof_mask == 0, kernel, (grid), (block), (smem), ctx->stream, __VA_ARGS__)); \
		ctx->launches++;                                                                                  \
	} while (0)

static const int MAX_PARTIALS = 1 << 20;
static int g_slab_ctx_on_device[64] = {0};   // live slab contexts of this process per device (see smd_slab_exchange_recv)
#ifndef SMD_PAIR3_MAX_DEFAULT
#define SMD_PAIR3_MAX_DEFAULT 32000   // measured (profiles/r02b_pair3_ab.md): 15 000 particles 72 -> 47 us per step, 30 000 76 -> 67, 60 000 78 -> 102
#endif
#ifndef SMD_DEFAULT_PDL_CHAIN
#define SMD_DEFAULT_PDL_CHAIN true
#endif

// particle count of a launch: by value on a single GPU, from the device word in slab mode (ctx->N is then only the
// launch bound).  cnt_ext: the count after an unpack (received particles appended, dead ghosts still in place).
static inline Cnt cnt_of(const smd_ctx *ctx) { Cnt c; c.n = ctx->N; c.dn = ctx->slab ? ctx->dN : nullptr; return c; }
static inline Cnt cnt_ext(const smd_ctx *ctx) { Cnt c; c.n = ctx->N; c.dn = ctx->slab ? ctx->dN + 1 : nullptr; return c; }

// ------------------------------------------------------------------------------------------------ phase timing
static cudaEvent_t prof_event(smd_ctx *ctx)
{
	cudaEvent_t e;
	if (!ctx->prof_free.empty()) { e = ctx->prof_free.back(); ctx->prof_free.pop_back(); return e; }
	cudaEventCreate(&e);
	return e;
}

static void prof_drain(smd_ctx *ctx)
{
	if (ctx->prof_pending.empty()) return;
	cudaEventSynchronize(ctx->prof_pending.back().e1);
	for (auto &sp : ctx->prof_pending) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) { ctx->prof_ms[sp.phase] += ms; ctx->prof_count[sp.phase]++; }
		ctx->prof_free.push_back(sp.e0);
		ctx->prof_free.push_back(sp.e1);
	}
	ctx->prof_pending.clear();
}

// NVTX ranges (tracing): SMD_NVTX=1 names every phase of the step on the host time line a profiler shows (Nsight
// Systems, ncu --nvtx); the ranges bracket the ENQUEUE of the phase's launches.  Header-only NVTX v3: without a tool attached
// the calls are a null function-pointer check.
static const char *const PHASE_NAMES[SMD_NPHASES] = {"smd:integrate1", "smd:build", "smd:pair", "smd:molecules", "smd:langevin", "smd:integrate2",
	"smd:step", "smd:exchange", "smd:fused_seam", "smd:pair+dU", "smd:build.hist", "smd:build.scan", "smd:build.place", "smd:build.reorder",
	"smd:phase14", "smd:phase15"};
static bool nvtx_on()
{
	static int v = -1;
	if (v < 0) { const char *e = getenv("SMD_NVTX"); v = (e && *e == '1') ? 1 : 0; }
	return v == 1;
}
struct NvtxRange {
	bool on;
	explicit NvtxRange(const char *name) : on(nvtx_on()) { if (on) nvtxRangePushA(name); }
	~NvtxRange() { if (on) nvtxRangePop(); }
};

// RAII bracket: records an event pair around the launches of one phase when that phase is enabled
struct ProfScope {
	smd_ctx *ctx; int phase; cudaEvent_t e0; bool on;
	NvtxRange nv;
	ProfScope(smd_ctx *c, int ph) : ctx(c), phase(ph), e0(nullptr), on((c->prof_mask >> ph) & 1u), nv(PHASE_NAMES[ph])
	{
		if (on) { e0 = prof_event(ctx); cudaEventRecord(e0, ctx->stream); }
	}
	~ProfScope()
	{
		if (!on) return;
		cudaEvent_t e1 = prof_event(ctx);
		cudaEventRecord(e1, ctx->stream);
		ctx->prof_pending.push_back({phase, e0, e1});
		if (ctx->prof_pending.size() >= 16384) prof_drain(ctx);
	}
};

extern "C" int smd_profile(smd_ctx *ctx, uint32_t phase_mask)
{
	if (!ctx) return SMD_ERR_ARG;
	cudaSetDevice(ctx->device);
	prof_drain(ctx);
	for (int p = 0; p < SMD_NPHASES; p++) { ctx->prof_ms[p] = 0; ctx->prof_count[p] = 0; }
	ctx->prof_mask = phase_mask;
	return SMD_OK;
}

extern "C" int smd_profile_read(smd_ctx *ctx, double ms[SMD_NPHASES], int64_t count[SMD_NPHASES])
{
	if (!ctx) return SMD_ERR_ARG;
	cudaSetDevice(ctx->device);
	prof_drain(ctx);
	for (int p = 0; p < SMD_NPHASES; p++) {
		if (ms) ms[p] = ctx->prof_ms[p];
		if (count) count[p] = ctx->prof_count[p];
	}
	return SMD_OK;
}

extern "C" int smd_abi_version(void) { return SMD_ABI_VERSION; }

extern "C" const char *smd_last_error(const smd_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int smd_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

// x slices per reference cell in the sort key (Geom::xs).  The sliced order no longer lists a cell's particles by
// descending original index, which is what the ORDERED look-up of an asymmetric constant table keys on (the reference's
// "later-loaded particle first", cellOpt.h:572-585): asymmetric tables keep xs = 1.
static void choose_xs(smd_ctx *ctx)
{
	Geom &g = ctx->geom;
	int xs = ctx->xs_wanted;
	if (ctx->tables_set && !ctx->tables_symmetric) xs = 1;
	const long long cap = ctx->cellcap > 0 ? ctx->cellcap : ctx->cellcap_limit;   // before / after the tables were allocated
	while (xs > 1 && ((long long)g.nc[0] * g.nc[1] * g.nc[2] * xs > cap)) xs >>= 1;
	if (xs < 1) xs = 1;
	if (g.xs != xs) ctx->cells_valid = false;
	g.xs = xs;
	g.finv = (double)xs / g.cs[0];
	ctx->pgeo.finv32 = (float)g.finv;
}

static void set_geom(smd_ctx *ctx, const double box[3])
{
	Geom &g = ctx->geom;
	for (int d = 0; d < 3; d++) {
		g.box[d] = box[d];
		g.nc[d] = (int)(box[d] / ctx->desc.cutoff);   // cellOpt.h:194-196 / :1529-1531
		g.cs[d] = box[d] / g.nc[d];                   // cellOpt.h:207-209 / :1560-1562
	}
	g.rc2 = ctx->desc.cutoff * ctx->desc.cutoff;
	g.slab = 0; g.col_lo = 0; g.col_hi = g.nc[0]; g.halo = 0;
	if (ctx->slab) {
		g.slab = 1;
		g.halo = SMD_SLAB_HALO;
		int32_t lo = 0, hi = g.nc[0];
		smd_slab_columns(g.nc[0], ctx->desc.nranks, ctx->desc.rank, &lo, &hi);
		g.col_lo = lo; g.col_hi = hi;
	}
	// FP32 prefilter threshold of k_pair_force2: absolute coordinates (and image shifts) up to maxL carry a rounding
	// error of maxL * 2^-24 each; a difference of two of them plus its own rounding stays below 4 of those, so
	// |r2_32 - r2_64| < 2 * sqrt(3) * rc * 4 * maxL * 2^-24 + (FP32 rounding of the squares).  32 * rc * that ulp
	// covers it several times over.
	double maxL = std::max(box[0], std::max(box[1], box[2]));
	double margin = 32.0 * ctx->desc.cutoff * maxL * (1.0 / 16777216.0) + 1e-5 * g.rc2;
	ctx->pgeo.thr32 = nextafterf((float)(g.rc2 + margin), INFINITY);
	ctx->pgeo.margin32 = nextafterf((float)margin, INFINITY);
	ctx->pgeo.slack32 = (float)(8.0 * maxL * (1.0 / 16777216.0) + 1e-6 * ctx->desc.cutoff);
	for (int d = 0; d < 3; d++) ctx->pgeo.cs32[d] = (float)g.cs[d];
	choose_xs(ctx);
}

// per-type phase-1 cutoff = raw + the current FP32 margin (changes with the box), capped at rc^2 + margin
static int pair_force_smem(smd_ctx *ctx, bool du);
static int drop_pending_histogram(smd_ctx *ctx);

static int upload_acut(smd_ctx *ctx)
{
	if (ctx->acut_raw.empty()) return SMD_OK;
	std::vector<float> &a = ctx->acut_host;   // lives in the context: the copy below is not waited for
	a.resize(ctx->nT);
	for (int t = 0; t < ctx->nT; t++)
		a[t] = ctx->acut_raw[t] < 0 ? -1.0f : std::min(ctx->acut_raw[t] + ctx->pgeo.margin32, ctx->pgeo.thr32);
	// a few bytes from pageable memory: staged by the runtime before the call returns, ordered on the stream
	cudaError_t e = cudaMemcpyAsync(ctx->acut, a.data(), a.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
	if (e != cudaSuccess) { ctx->err = std::string("upload_acut: ") + cudaGetErrorString(e); return SMD_ERR_CUDA; }
	return SMD_OK;
}

static int check_geom(smd_ctx *ctx)
{
	const Geom &g = ctx->geom;
	for (int d = 0; d < 3; d++) {
		if (g.nc[d] < 3) {
			ctx->err = "box shorter than 3 cells along an axis: the reference's half stencil double-counts there; unsupported";
			return SMD_ERR_UNSUPPORTED;
		}
	}
	if (g.nc[0] > 2047 || g.nc[1] > 2047 || g.nc[2] > 1023) {
		ctx->err = "cell grid exceeds 2047 x 2047 x 1023 cells";
		return SMD_ERR_UNSUPPORTED;
	}
	if (ctx->slab && g.col_hi - g.col_lo < 2 * g.halo + 1) {
		ctx->err = "slab narrower than 2 * halo + 1 cell columns: use fewer ranks for this box";
		return SMD_ERR_UNSUPPORTED;
	}
	return SMD_OK;
}

extern "C" int smd_create(const smd_desc *desc, smd_ctx **out)
{
	if (!desc || !out) { g_create_error = "null argument"; return SMD_ERR_ARG; }
	if (desc->abi_version != SMD_ABI_VERSION) { g_create_error = "ABI version mismatch"; return SMD_ERR_ARG; }
	if (desc->n_particles < 0 || desc->n_types <= 0 || desc->cutoff <= 0 || desc->dt <= 0) {
		g_create_error = "invalid descriptor (n_particles, n_types, cutoff, dt)";
		return SMD_ERR_ARG;
	}
	int ndev = smd_device_count();
	if (ndev <= 0 || desc->device < 0 || desc->device >= ndev) {
		g_create_error = "no usable CUDA device (this library has no CPU fallback)";
		return SMD_ERR_CUDA;
	}
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, desc->device) != cudaSuccess || prop.major < 10) {
		g_create_error = "device is not sm_100 class";
		return SMD_ERR_CUDA;
	}
	smd_ctx *ctx = new smd_ctx();
	ctx->desc = *desc;
	ctx->N = desc->n_particles;
	ctx->n_global = desc->n_particles;
	ctx->nT = desc->n_types;
	ctx->cap = ((std::max(ctx->N, 1) + 255) / 256) * 256;
	ctx->slab = desc->nranks > 1;
	if (ctx->slab) g_slab_ctx_on_device[desc->device & 63]++;
	memset(&ctx->comm, 0, sizeof ctx->comm);
	if (ctx->slab) {
		if (desc->rank < 0 || desc->rank >= desc->nranks) { g_create_error = "rank outside [0, nranks)"; delete ctx; return SMD_ERR_ARG; }
		if (desc->n_particles >= GID_GHOST) { g_create_error = "slab mode supports fewer than 2^30 particles"; delete ctx; return SMD_ERR_UNSUPPORTED; }
		if (desc->noise == SMD_NOISE_EXTERNAL) { g_create_error = "slab mode uses the Philox noise only"; delete ctx; return SMD_ERR_UNSUPPORTED; }
		long long want = desc->reserved[0] > 0 ? desc->reserved[0]
		                                       : (long long)(1.25 * desc->n_particles / desc->nranks) + 65536;
		ctx->cap = (int)(((std::min<long long>(want, (long long)desc->n_particles + 65536) + 255) / 256) * 256);
		ctx->N = ctx->cap;   // launch bound; the live count is the device word dN[0]
		ctx->comm.capmsg = desc->reserved[1] > 0 ? desc->reserved[1] : std::max(16384, ctx->cap / 8);
	}
	ctx->temperature = desc->temperature;
	ctx->device = desc->device;
	ctx->cur = 0;
	ctx->pcur = 0;
	ctx->cells_valid = false;
	ctx->tables_set = ctx->particles_set = false;
	ctx->noise_ready = false;
	ctx->acc_live = false;
	ctx->n_molecules = 0;
	ctx->launches = ctx->rebuilds = 0;
	{ const char *e = getenv("SMD_ENERGY_ONEPHASE"); ctx->force_onephase_energy = e && *e == '1'; }
	// SMD_NO_FUSE=1: every phase of the step in its own kernel (A/B and bit-identity tests).
	{ const char *e = getenv("SMD_NO_FUSE"); ctx->no_fuse = e && *e == '1'; }
	{ const char *e = getenv("SMD_PAIR3"); ctx->pair3 = e ? atoi(e) : -1; }
	{ const char *e = getenv("SMD_PAIR3_MAX"); ctx->pair3_max = e ? atoi(e) : SMD_PAIR3_MAX_DEFAULT; }
	{ const char *e = getenv("SMD_PAIR_TAIL"); ctx->pair_tail = (e && *e == '1') ? 1 : 0; }
	{ cudaDeviceProp prop; if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.multiProcessorCount > 0) ctx->sm_count = prop.multiProcessorCount; }
	{ const char *e = getenv("SMD_NO_PREBIN"); ctx->prebin = !(e && *e == '1'); }
	{ const char *e = getenv("SMD_NO_SLAB_PREBIN"); ctx->no_slab_prebin = e && *e == '1'; }
	{ const char *e = getenv("SMD_NO_DU_FUSE"); ctx->no_du_fuse = e && *e == '1'; }
	{ const char *e = getenv("SMD_NO_SEAM_PACK"); ctx->no_seam_pack = e && *e == '1'; }
	// SMD_PDL=0: plain stream order everywhere; 1: only the step seam is a programmatic dependent (of the pair kernel);
	// 2 (default): so are the kernels of the build, the pair kernel itself and the slab unpack (LAUNCHP).  Measured per MD
	// step: C2 218.3 / 213.9 / 203.4 us, a 15 000-particle vesicle 72.7 / 77.4 / 66.3 us.
	{ const char *e = getenv("SMD_PDL"); ctx->pdl = !(e && *e == '0'); ctx->pdl_chain = e ? (*e == '2') : SMD_DEFAULT_PDL_CHAIN; }   // slab mode: exchange packed by a kernel of its own (A/B)   // smd_step_mc: dPotential in a pass of its own (A/B)
	{ const char *e = getenv("SMD_XSUB"); int v = e ? atoi(e) : 4; ctx->xs_wanted = (v == 1 || v == 2 || v == 4 || v == 8) ? v : 4; }
	ctx->pcur = 0;
	set_geom(ctx, desc->box);
	ctx->pgeo.rmin32 = (float)desc->cutoff;
	ctx->pgeo.done = nullptr; ctx->pgeo.epoch = 0; ctx->pgeo.first = 0; ctx->pgeo.part0 = 0; ctx->pgeo.nowait = 0;
	int rc = check_geom(ctx);
	if (rc) { g_create_error = ctx->err; delete ctx; return rc; }

#define CKC(call)                                                                                         \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) {                                                                          \
			g_create_error = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
			delete ctx;                                                                                   \
			return SMD_ERR_CUDA;                                                                          \
		}                                                                                                 \
	} while (0)
	CKC(cudaSetDevice(ctx->device));
	CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	size_t cap = ctx->cap;
	for (int b = 0; b < 2; b++) {
		CKC(cudaMalloc(&ctx->pos[b], cap * sizeof(Particle)));
		CKC(cudaMalloc(&ctx->vel[b], 3 * cap * sizeof(double)));
		CKC(cudaMalloc(&ctx->gid[b], cap * sizeof(int)));
		ctx->unw[b] = nullptr;
		if (desc->track_unwrapped) CKC(cudaMalloc(&ctx->unw[b], 3 * cap * sizeof(double)));
	}
	CKC(cudaMalloc(&ctx->pos32, (cap + 8) * sizeof(float4)));   // + overhang of the 4-wide candidate loads and their prefetch
	CKC(cudaMemset(ctx->pos32, 0, (cap + 8) * sizeof(float4)));
	CKC(cudaMalloc(&ctx->acut, (size_t)ctx->nT * sizeof(float)));
	CKC(cudaMalloc(&ctx->pos16, (cap + 16) * sizeof(uint2)));   // + overhang, as for pos32
	CKC(cudaMemset(ctx->pos16, 0, (cap + 16) * sizeof(uint2)));
	CKC(cudaMalloc(&ctx->arad, (size_t)ctx->nT * sizeof(float)));
	CKC(cudaMalloc(&ctx->ptab, (size_t)PTAB_STRIDE * ctx->nT * ctx->nT * sizeof(double)));
	CKC(cudaMalloc(&ctx->utab, (size_t)PTAB_STRIDE * ctx->nT * ctx->nT * sizeof(double)));
	CKC(cudaMalloc(&ctx->win, 2 * WIN_WORDS * sizeof(int)));
	CKC(cudaMemset(ctx->win, 0, 2 * WIN_WORDS * sizeof(int)));
	CKC(cudaMalloc(&ctx->acc, 3 * cap * sizeof(double)));
	CKC(cudaMemset(ctx->acc, 0, 3 * cap * sizeof(double)));
	CKC(cudaMalloc(&ctx->acc2, 3 * cap * sizeof(double)));
	CKC(cudaMalloc(&ctx->slot_of, (size_t)std::max<size_t>(cap, (size_t)ctx->n_global) * sizeof(int)));
	CKC(cudaMemset(ctx->slot_of, 0xff, (size_t)std::max<size_t>(cap, (size_t)ctx->n_global) * sizeof(int)));
	CKC(cudaMalloc(&ctx->dN, 2 * sizeof(int)));
	CKC(cudaMemset(ctx->dN, 0, 2 * sizeof(int)));
	CKC(cudaMalloc(&ctx->d_export_counter, sizeof(int)));
	for (int b = 0; b < 2; b++) CKC(cudaMemset(ctx->gid[b], 0, cap * sizeof(int)));
	if (ctx->slab) {
		// receive buffers: [side][parity] = header + capmsg entries; the neighbours write them through peer memory
		ctx->comm.parity_stride = (sizeof(SlabMsgHeader) + (size_t)ctx->comm.capmsg * sizeof(SlabMsgEntry) + 255) / 256 * 256;
		ctx->recv_bytes = 2 * ctx->comm.parity_stride;
		for (int sd = 0; sd < 2; sd++) {
			CKC(cudaMalloc(&ctx->recv_base[sd], ctx->recv_bytes));
			CKC(cudaMemset(ctx->recv_base[sd], 0, ctx->recv_bytes));
			ctx->comm.recv[sd] = ctx->recv_base[sd];
		}
		{ const char *e = getenv("SMD_SLAB_PULL"); ctx->comm.pull = (e && *e == '1') ? 1 : 0; }
		{ const char *e = getenv("SMD_SLAB_FENCE_GPU"); ctx->comm.gpu_fence = (e && *e == '1') ? 1 : 0; }
		CKC(cudaMalloc(&ctx->comm.counters, 4 * sizeof(int)));
		CKC(cudaMemset(ctx->comm.counters, 0, 4 * sizeof(int)));
	}
	// dense offset table over the occupied window of the reference grid; capacity = whole grid up to 64 Mi cells
	long long total = (long long)ctx->geom.nc[0] * ctx->geom.nc[1] * ctx->geom.nc[2];
	ctx->cellcap = std::min<long long>(std::max<long long>(std::max<long long>(2, ctx->geom.xs) * total, 1 << 16), 64ll << 20);
	CKC(cudaMalloc(&ctx->count, (ctx->cellcap + 1) * sizeof(int)));
	CKC(cudaMemset(ctx->count, 0, (ctx->cellcap + 1) * sizeof(int)));
	CKC(cudaMalloc(&ctx->start, (ctx->cellcap + 1) * sizeof(int)));
	CKC(cudaMalloc(&ctx->cursor, (ctx->cellcap + 1) * sizeof(int)));
	CKC(cudaMalloc(&ctx->scan_state, (size_t)(2 * SCAN_BLOCKS) * sizeof(unsigned long long)));
	CKC(cudaMemset(ctx->scan_state, 0, (size_t)(2 * SCAN_BLOCKS) * sizeof(unsigned long long)));
	CKC(cudaMalloc(&ctx->scan_barrier, 2 * sizeof(unsigned)));   // [0] grid barrier of the fallback, [1] arrival tickets
	CKC(cudaMemset(ctx->scan_barrier, 0, 2 * sizeof(unsigned)));
	CKC(cudaMalloc(&ctx->cellOfSlot, cap * sizeof(int)));
	CKC(cudaMalloc(&ctx->order, cap * sizeof(int2)));
	CKC(cudaMalloc(&ctx->bbox, 6 * sizeof(int)));
	int bb[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
	CKC(cudaMemcpy(ctx->bbox, bb, sizeof bb, cudaMemcpyHostToDevice));
	CKC(cudaMalloc(&ctx->errflag, sizeof(int)));
	CKC(cudaMemset(ctx->errflag, 0, sizeof(int)));
	CKC(cudaMalloc(&ctx->fC, 6 * ctx->nT * ctx->nT * sizeof(double)));
	CKC(cudaMalloc(&ctx->uC, 6 * ctx->nT * ctx->nT * sizeof(double)));
	CKC(cudaMalloc(&ctx->noise, 3 * cap * sizeof(double)));
	CKC(cudaMalloc(&ctx->partials, MAX_PARTIALS * sizeof(double)));
	CKC(cudaMalloc(&ctx->scalars, 64 * sizeof(double)));
	CKC(cudaMalloc(&ctx->icount, cap * sizeof(int)));
	CKC(cudaMalloc(&ctx->stage, 3 * cap * sizeof(double) * 2));
	CKC(cudaMalloc(&ctx->istage, 2 * cap * sizeof(int)));
	CKC(cudaMallocHost(&ctx->h_pinned, 64 * sizeof(double)));
#undef CKC
	{
		// the opt-in shared-memory size is a property of the FUNCTION on this device, shared by every context of the
		// process: only ever raise it (a later context with fewer particle types must not shrink it under an earlier one)
		static int smem_set[64] = {0};
		int smem = pair_force_smem(ctx, false);
		int &have = smem_set[ctx->device & 63];
		cudaError_t e1 = cudaSuccess;
		if (smem > have) {
			e1 = cudaFuncSetAttribute(k_pair_force2<0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
			if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_pair_force2<0, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
			if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_pair_force2<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
			if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_pair_force2<0, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
			if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_pair_force2<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
			if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_pair_force2<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
			if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_pair_force2<3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair_force_smem(ctx, true));
			if (e1 == cudaSuccess) have = smem;
		}
		if (e1 != cudaSuccess) {
			g_create_error = std::string("cudaFuncSetAttribute(k_pair_force2): ") + cudaGetErrorString(e1);
			smd_destroy(ctx);
			return SMD_ERR_CUDA;
		}
	}
	if (6 * ctx->nT * ctx->nT * sizeof(double) > 40000) {
		g_create_error = "too many particle types for the shared-memory pair table";
		smd_destroy(ctx);
		return SMD_ERR_UNSUPPORTED;
	}
	*out = ctx;
	return SMD_OK;
}

extern "C" int smd_destroy(smd_ctx *ctx)
{
	if (!ctx) return SMD_OK;
	if (ctx->slab) g_slab_ctx_on_device[ctx->device & 63]--;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for (int b = 0; b < 2; b++) {
		cudaFree(ctx->pos[b]); cudaFree(ctx->vel[b]); cudaFree(ctx->gid[b]); cudaFree(ctx->unw[b]);
	}
	cudaFree(ctx->pos32); cudaFree(ctx->pos16); cudaFree(ctx->arad); cudaFree(ctx->acut); cudaFree(ctx->ptab); cudaFree(ctx->utab);
	cudaFree(ctx->win); cudaFree(ctx->acc); cudaFree(ctx->acc2); cudaFree(ctx->slot_of); cudaFree(ctx->count); cudaFree(ctx->start); cudaFree(ctx->cursor);
	cudaFree(ctx->scan_state); cudaFree(ctx->scan_barrier); cudaFree(ctx->cellOfSlot); cudaFree(ctx->order); cudaFree(ctx->bbox); cudaFree(ctx->errflag);
	cudaFree(ctx->fC); cudaFree(ctx->uC); cudaFree(ctx->noise); cudaFree(ctx->partials); cudaFree(ctx->scalars);
	cudaFree(ctx->icount); cudaFree(ctx->stage); cudaFree(ctx->istage);
	cudaFreeHost(ctx->h_pinned);
	cudaFree(ctx->dN); cudaFree(ctx->d_export_counter);
	for (int sd = 0; sd < 2; sd++) {
		if (ctx->ipc_opened[sd]) cudaIpcCloseMemHandle(ctx->ipc_opened[sd]);
		cudaFree(ctx->recv_base[sd]);
	}
	cudaFree(ctx->comm.counters);
	if (ctx->snap_stage) {
		cudaFree(ctx->snap_stage); cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->snap_gathered);
		cudaEventDestroy(ctx->snap_done[0]); cudaEventDestroy(ctx->snap_done[1]);
	}
	for (auto &b : ctx->bonds) cudaFree(b.d_ij);
	for (auto &b : ctx->bends) cudaFree(b.d_ijk);
	for (auto &b : ctx->balls) cudaFree(b.d_cj);
	for (auto &b : ctx->beads) { cudaFree(b.d_beads); cudaFree(b.d_C); }
	for (auto &f : ctx->fields) { cudaFree(f.d_idx); cudaFree(f.d_C); }
	prof_drain(ctx);
	for (auto e : ctx->prof_free) cudaEventDestroy(e);
	if (ctx->pair_done) cudaFree(ctx->pair_done);
	if (ctx->du_partials) cudaFree(ctx->du_partials);
	if (ctx->terms_dev) cudaFree(ctx->terms_dev);
	if (ctx->import_bad) cudaFree(ctx->import_bad);
	if (ctx->export_i) cudaFree(ctx->export_i);
	if (ctx->export_d) cudaFree(ctx->export_d);
	if (ctx->obs_buf) cudaFree(ctx->obs_buf);
	if (ctx->obs_host) cudaFreeHost(ctx->obs_host);
	if (ctx->ke_bins) cudaFree(ctx->ke_bins);
	if (ctx->ke_overflow) cudaFree(ctx->ke_overflow);
	if (ctx->msd_start) cudaFree(ctx->msd_start);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
	return SMD_OK;
}

static int check_device_errors(smd_ctx *ctx)
{
	int flag = 0;
	CK(cudaMemcpyAsync(&flag, ctx->errflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaGetLastError());
	if (flag) {
		ctx->err = "device-side cell error:";
		if (flag & ERR_OUT_OF_BOX) ctx->err += " particle outside the box or NaN position (reference: 'cell placement is on boundary', cellOpt.h:541-552)";
		if (flag & ERR_WINDOW_CAP) ctx->err += " occupied cell window exceeds the offset-table capacity";
		if (flag & ERR_SLAB_MIGRATION) ctx->err += " slab: a particle left its slab by more than the halo width in one step";
		if (flag & ERR_SLAB_MSG_CAP) ctx->err += " slab: halo message capacity exceeded (desc.reserved[1])";
		if (flag & ERR_SLAB_CAPACITY) ctx->err += " slab: local particle capacity exceeded (desc.reserved[0])";
		if (flag & ERR_SLAB_TIMEOUT) ctx->err += " slab: timed out waiting for a neighbour's halo message";
		if (flag & ERR_SLAB_MISSING) ctx->err += " slab: a member of a chain / bond / bend record is neither owned nor inside the halo";
		return SMD_ERR_CELL;
	}
	return SMD_OK;
}

extern "C" int smd_synchronize(smd_ctx *ctx)
{
	if (!ctx) return SMD_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	return check_device_errors(ctx);
}

extern "C" int smd_set_pair_tables(smd_ctx *ctx, const double *fC, const double *uC)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(fC && uC, "null table");
	CK(cudaSetDevice(ctx->device));
	{ int rcd = drop_pending_histogram(ctx); if (rcd) return rcd; }   // (the sort key may change with the tables: choose_xs)
	size_t bytes = 6 * (size_t)ctx->nT * ctx->nT * sizeof(double);
	CK(cudaMemcpyAsync(ctx->fC, fC, bytes, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(ctx->uC, uC, bytes, cudaMemcpyHostToDevice, ctx->stream));
	{
		// phase-1 class cutoff per type and the padded phase-2 table of k_pair_force2 (see there)
		int nT = ctx->nT;
		auto row = [&](int a, int b) { return fC + 6 * (a * nT + b); };
		auto zero_tail = [&](const double *r) { return r[4] == 0.0 && r[5] == 0.0 && ctx->desc.cutoff <= 2.0 * r[0] && r[0] > 0; };
		auto zero_core = [&](const double *r) { return r[1] == 0.0 && r[2] == 0.0; };
		ctx->acut_raw.assign(nT, -1.0f);
		for (int a = 0; a < nT; a++)
			for (int b = 0; b < nT; b++) {
				const double *r1 = row(a, b), *r2 = row(b, a);   // looked up in either orientation
				float c = INFINITY;
				if (zero_tail(r1) && zero_tail(r2)) {
					double rm = std::max(r1[0], r2[0]);
					c = nextafterf((float)(rm * rm), INFINITY);
					if (zero_core(r1) && zero_core(r2)) c = -1.0f;
				}
				ctx->acut_raw[a] = std::max(ctx->acut_raw[a], c);
			}
		{   // class radius per type for the 16-bit phase-1 records: rm for purely repulsive types, rc else, -1: none
			std::vector<float> rad(nT);
			float rmin = (float)ctx->desc.cutoff;
			for (int t = 0; t < nT; t++) {
				float raw = ctx->acut_raw[t];
				rad[t] = raw < 0 ? -1.0f : (std::isinf(raw) ? (float)ctx->desc.cutoff : std::min(nextafterf(sqrtf(raw), INFINITY), (float)ctx->desc.cutoff));
				if (rad[t] > 0) rmin = std::min(rmin, rad[t]);
			}
			ctx->pgeo.rmin32 = rmin;
			CK(cudaMemcpyAsync(ctx->arad, rad.data(), rad.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
			CK(cudaStreamSynchronize(ctx->stream));
		}
		// smallest double whose correctly rounded square root is >= v
		auto sq_threshold = [](double v) {
			double x = v * v;
			while (x > 0 && sqrt(x) >= v) x = nextafter(x, 0.0);
			while (sqrt(x) < v) x = nextafter(x, INFINITY);
			return x;
		};
		std::vector<double> pt((size_t)PTAB_STRIDE * nT * nT, 0.0);
		for (int k = 0; k < nT * nT; k++) {
			const double *r = fC + 6 * k;
			double *o = pt.data() + (size_t)PTAB_STRIDE * k;
			o[0] = r[0] > 0 ? sq_threshold(r[0]) : 0.0;
			o[1] = r[0] > 0 ? sq_threshold(r[0] + r[0]) : 0.0;
			o[2] = r[0]; o[3] = r[1]; o[4] = r[2];
			o[6] = r[3]; o[7] = r[4]; o[8] = r[5];
		}
		CK(cudaMemcpyAsync(ctx->ptab, pt.data(), pt.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		// the potential's constants in the same padded layout (energy modes of the two-phase kernel)
		std::vector<double> ut((size_t)PTAB_STRIDE * nT * nT, 0.0);
		for (int k = 0; k < nT * nT; k++) {
			const double *r = uC + 6 * k;
			double *o = ut.data() + (size_t)PTAB_STRIDE * k;
			o[2] = r[0]; o[3] = r[1]; o[4] = r[2];
			o[6] = r[3]; o[7] = r[4]; o[8] = r[5];
		}
		CK(cudaMemcpyAsync(ctx->utab, ut.data(), ut.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		int rc = upload_acut(ctx);
		if (rc) return rc;
	}
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->tables_set = true;
	// the reference looks rows up in the order (home particle, neighbour); every shipped generator
	// writes symmetric tables, which lets the pair kernel skip the orientation logic.  Asymmetric tables still work.
	ctx->tables_symmetric = true;
	for (int a = 0; a < ctx->nT; a++)
		for (int b = 0; b < a; b++)
			for (int k = 0; k < 6; k++)
				if (fC[6 * (a * ctx->nT + b) + k] != fC[6 * (b * ctx->nT + a) + k] || uC[6 * (a * ctx->nT + b) + k] != uC[6 * (b * ctx->nT + a) + k])
					ctx->tables_symmetric = false;
	choose_xs(ctx);
	return SMD_OK;
}

// tag every particle with its cell and accumulate the occupied bounding box (after set_particles / a box move;
// the steady state does this inside k_verlet_first)
static int retag_cells(smd_ctx *ctx, bool rearm = true)
{
	// rearm = false: the cell grid is the same and the particles barely moved (accepted box move): keep the extremes,
	// so that the tagging pass issues almost no atomics (thousands of warps hitting one address cost ~40 us)
	int rcd = drop_pending_histogram(ctx);
	if (rcd) return rcd;
	if (rearm) LAUNCH(k_arm_bbox, 1, 32, 0, ctx->bbox);
	LAUNCH(k_tag_cells, nblk(ctx->N, TPB), TPB, 0, cnt_of(ctx), ctx->pos[ctx->pcur], ctx->geom, ctx->bbox, ctx->errflag, ctx->gid[ctx->cur]);
	ctx->cells_valid = false;
	return SMD_OK;
}

extern "C" int smd_set_particles(smd_ctx *ctx, const double *xyz, const int32_t *type, const double *vel)
{
	if (!ctx) return SMD_ERR_ARG;
	ctx->du_ready = false;   // the particles change: block sums armed by smd_arm_dpotential no longer describe them
	REQUIRE(xyz && type, "null positions / types");
	CK(cudaSetDevice(ctx->device));
	int N = ctx->n_global;
	// the reference refuses out-of-box particles at load (system.h:452-469).  One GPU: checked by the import kernel (0.7 ms of
	// host loop per 240 000 particles otherwise, on the end-to-end path of every upload); slab mode walks the arrays anyway.
	const bool device_check = !ctx->slab;
	for (int i = 0; i < N && !device_check; i++)
		for (int d = 0; d < 3; d++) {
			double x = xyz[3 * i + d];
			if (!(x >= 0 && x <= ctx->geom.box[d])) {
				char buf[128];
				snprintf(buf, sizeof buf, "%c position of particle %d is out of bounds.", "XYZ"[d], i);
				ctx->err = buf;
				return SMD_ERR_CELL;
			}
		}
	for (int i = 0; i < N && !device_check; i++)
		if (type[i] < 0 || type[i] >= ctx->nT) { ctx->err = "particle type out of range"; return SMD_ERR_ARG; }
	double *sx = ctx->stage, *sv = ctx->stage + 3 * (size_t)ctx->cap;
	const int *gid_in = nullptr;
	std::vector<double> lx, lv;
	std::vector<int> lt, lg;
	if (ctx->slab) {
		// keep what this rank owns plus its ghost columns; every rank was handed the same global arrays
		std::vector<int32_t> flags((size_t)N);
		smd_slab_select(ctx->geom.box, ctx->desc.cutoff, ctx->desc.nranks, ctx->desc.rank, N, xyz, flags.data());
		for (int i = 0; i < N; i++) {
			if (!flags[i]) continue;
			for (int d = 0; d < 3; d++) { lx.push_back(xyz[3 * i + d]); lv.push_back(vel ? vel[3 * i + d] : 0.0); }
			lt.push_back(type[i]);
			lg.push_back(flags[i] == 2 ? (i | GID_GHOST) : i);
		}
		N = (int)lt.size();
		if (N > ctx->cap) { ctx->err = "slab: local particle capacity exceeded at load (desc.reserved[0])"; return SMD_ERR_ARG; }
		xyz = lx.data(); type = lt.data(); vel = lv.data();
		int *dg = ctx->istage + ctx->cap;
		CK(cudaMemcpyAsync(dg, lg.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
		gid_in = dg;
		CK(cudaMemsetAsync(ctx->slot_of, 0xff, (size_t)ctx->n_global * sizeof(int), ctx->stream));
		int nn[2] = {N, N};
		CK(cudaMemcpyAsync(ctx->dN, nn, sizeof nn, cudaMemcpyHostToDevice, ctx->stream));
		ctx->exch_pending = false;
	}
	CK(cudaMemcpyAsync(sx, xyz, 3 * (size_t)N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(ctx->istage, type, (size_t)N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	if (vel) CK(cudaMemcpyAsync(sv, vel, 3 * (size_t)N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	ctx->cur = 0;
	ctx->pcur = 0;
	int *bad = nullptr, *h_bad = reinterpret_cast<int *>(ctx->h_pinned + 60);
	if (device_check) {
		if (!ctx->import_bad) CK(cudaMalloc(&ctx->import_bad, 2 * sizeof(int)));
		bad = ctx->import_bad;
		CK(cudaMemsetAsync(bad, 0x7f, 2 * sizeof(int), ctx->stream));   // 0x7f7f7f7f: larger than any code
	}
	if (N > 0)
		LAUNCH(k_import_particles, nblk(N, TPB), TPB, 0, N, ctx->cap, sx, ctx->istage, vel ? sv : nullptr, ctx->pos[0], ctx->vel[0],
		       ctx->unw[0], ctx->gid[0], ctx->slot_of, gid_in, ctx->geom, ctx->nT, bad, ctx->n_global);
	if (bad) CK(cudaMemcpyAsync(h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaMemsetAsync(ctx->acc, 0, 3 * (size_t)ctx->cap * sizeof(double), ctx->stream));
	ctx->acc_live = false;
	retag_cells(ctx);
	CK(cudaStreamSynchronize(ctx->stream));
	if (bad && *h_bad != 0x7f7f7f7f) {
		const int i = *h_bad / 4, d = *h_bad % 4;
		ctx->particles_set = false;
		CK(cudaMemsetAsync(ctx->errflag, 0, sizeof(int), ctx->stream));   // (the tagging pass saw the same particle: that report is this one)
		if (d == 3) { ctx->err = "particle type out of range"; return SMD_ERR_ARG; }
		char buf[128];
		snprintf(buf, sizeof buf, "%c position of particle %d is out of bounds.", "XYZ"[d], i);
		ctx->err = buf;
		return SMD_ERR_CELL;
	}
	ctx->particles_set = true;
	ctx->noise_ready = false;
	return SMD_OK;
}

static int check_index(smd_ctx *ctx, const int32_t *v, size_t n, const char *what)
{
	for (size_t i = 0; i < n; i++)
		if (v[i] < 0 || v[i] >= ctx->n_global) {
			ctx->err = std::string(what) + " index out of bounds";   // Blob::errorChecking, system.h:483-545
			return SMD_ERR_ARG;
		}
	return SMD_OK;
}

extern "C" int smd_add_chain(smd_ctx *ctx, int32_t n_blocks, const int32_t *blocks, const double c[4])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n_blocks >= 0 && (blocks || n_blocks == 0) && c, "bad CHAIN arguments");
	for (int j = 0; j < n_blocks; j++) {
		ChainBlock cb;
		cb.start = blocks[3 * j]; cb.nChains = blocks[3 * j + 1]; cb.len = blocks[3 * j + 2];
		long long end = (long long)cb.start + (long long)cb.nChains * cb.len;
		REQUIRE(cb.start >= 0 && cb.nChains >= 0 && end <= ctx->n_global, "CHAIN Molecule is out of bounds!");
		REQUIRE(cb.len >= 3, "CHAIN length below 3 is undefined in the reference (system.h:1834-1836)");
		for (int k = 0; k < 4; k++) cb.c[k] = c[k];
		ctx->chains.push_back(cb);
	}
	ctx->mol_order.push_back({SMD_MOL_CHAIN, (int)ctx->chains.size() - n_blocks, n_blocks});
	ctx->n_molecules++;
	return SMD_OK;
}

template <class L>
static int upload_list(smd_ctx *ctx, const int32_t *v, size_t n, int **dptr)
{
	*dptr = nullptr;
	if (n == 0) return SMD_OK;
	CK(cudaSetDevice(ctx->device));
	CK(cudaMalloc(dptr, n * sizeof(int)));
	CK(cudaMemcpy(*dptr, v, n * sizeof(int), cudaMemcpyHostToDevice));
	return SMD_OK;
}

extern "C" int smd_add_bonds(smd_ctx *ctx, int32_t n, const int32_t *ij, const double c[2])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n >= 0 && (ij || n == 0) && c, "bad BOND arguments");
	int rc = check_index(ctx, ij, 2 * (size_t)n, "BOND Molecule");
	if (rc) return rc;
	BondList b;
	b.n = n; b.c[0] = c[0]; b.c[1] = c[1];
	rc = upload_list<int>(ctx, ij, 2 * (size_t)n, &b.d_ij);
	if (rc) return rc;
	ctx->bonds.push_back(b);
	ctx->mol_order.push_back({SMD_MOL_BOND, (int)ctx->bonds.size() - 1, 1});
	ctx->n_molecules++;
	return SMD_OK;
}

extern "C" int smd_add_bends(smd_ctx *ctx, int32_t n, const int32_t *ijk, const double c[2])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n >= 0 && (ijk || n == 0) && c, "bad BEND arguments");
	int rc = check_index(ctx, ijk, 3 * (size_t)n, "BEND Molecule");
	if (rc) return rc;
	BendList b;
	b.n = n; b.c[0] = c[0]; b.c[1] = c[1];
	rc = upload_list<int>(ctx, ijk, 3 * (size_t)n, &b.d_ijk);
	if (rc) return rc;
	ctx->bends.push_back(b);
	ctx->mol_order.push_back({SMD_MOL_BEND, (int)ctx->bends.size() - 1, 1});
	ctx->n_molecules++;
	return SMD_OK;
}

extern "C" int smd_add_ball(smd_ctx *ctx, int32_t n, const int32_t *cj, const double c[2])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n >= 0 && (cj || n == 0) && c, "bad BALL arguments");
	if (ctx->slab) { ctx->err = "slab mode supports CHAIN, BOND and BEND molecules only"; return SMD_ERR_UNSUPPORTED; }
	int rc = check_index(ctx, cj, 2 * (size_t)n, "BALL Molecule");
	if (rc) return rc;
	BallList b;
	b.n = n; b.c[0] = c[0]; b.c[1] = c[1];
	rc = upload_list<int>(ctx, cj, 2 * (size_t)n, &b.d_cj);
	if (rc) return rc;
	ctx->balls.push_back(b);
	ctx->mol_order.push_back({SMD_MOL_BALL, (int)ctx->balls.size() - 1, 1});
	ctx->n_molecules++;
	return SMD_OK;
}

// ---- the remaining molecule kinds of MD.cpp's switch (MD.cpp:414-478): one-body fields and NANOCORE
static int add_field(smd_ctx *ctx, int kind, int32_t n, const int32_t *idx, const double *c, int nc_host, const double *C, size_t nC,
                     const int32_t *blocks, int width, const char *what)
{
	if (ctx->slab) { ctx->err = "slab mode supports CHAIN, BOND and BEND molecules only"; return SMD_ERR_UNSUPPORTED; }
	CK(cudaSetDevice(ctx->device));
	FieldMol f;
	f.kind = kind; f.n = n; f.d_idx = nullptr; f.d_C = nullptr;
	for (int k = 0; k < 4; k++) f.c[k] = (c && k < nc_host) ? c[k] : 0.0;
	if (idx) {
		int rc = check_index(ctx, idx, (size_t)n, what);
		if (rc) return rc;
		rc = upload_list<int>(ctx, idx, (size_t)n, &f.d_idx);
		if (rc) return rc;
	}
	if (blocks) f.blocks.assign(blocks, blocks + (size_t)n * width);
	if (kind == SMD_MOL_NANOCORE)
		for (int j = 0; j < n; j++) { f.host_idx.push_back(idx[j]); f.host_radius.push_back(C[22 * j + 4]); }   // BEADRADIUS, MD.h:53
	if (C && nC) {
		CK(cudaMalloc(&f.d_C, nC * sizeof(double)));
		CK(cudaMemcpy(f.d_C, C, nC * sizeof(double), cudaMemcpyHostToDevice));
	}
	ctx->fields.push_back(f);
	ctx->mol_order.push_back({kind, (int)ctx->fields.size() - 1, 1});
	ctx->n_molecules++;
	return SMD_OK;
}

extern "C" int smd_add_boundary(smd_ctx *ctx, int32_t n, const int32_t *idx, const double c[4])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n >= 0 && (idx || n == 0) && c, "bad BOUNDARY arguments");
	int dim = (int)c[0];   // static_cast<int>(constants[0]), MD.h:569
	REQUIRE(dim >= 0 && dim <= 2, "BOUNDARY: constants[0] must name an axis (0, 1 or 2)");
	return add_field(ctx, SMD_MOL_BOUNDARY, n, idx, c, 4, nullptr, 0, nullptr, 0, "BOUNDARY Molecule");
}

extern "C" int smd_add_floating_base(smd_ctx *ctx, int32_t n, const int32_t *idx, const double *C)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n >= 0 && (idx || n == 0) && C, "bad FLOATING_BASE arguments");
	return add_field(ctx, SMD_MOL_FLOATING_BASE, n, idx, nullptr, 0, C, 6 * (size_t)ctx->nT, nullptr, 0, "FLOATING_BASE Molecule");
}

extern "C" int smd_add_ztorque(smd_ctx *ctx, int32_t n_blocks, const int32_t *blocks, const double c[4])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n_blocks >= 0 && (blocks || n_blocks == 0) && c, "bad ZTORQUE arguments");
	for (int j = 0; j < n_blocks; j++) {
		long long start = blocks[3 * j], nCh = blocks[3 * j + 1], len = blocks[3 * j + 2];
		REQUIRE(start >= 0 && nCh >= 0 && len >= 0 && start + nCh * len <= ctx->n_global, "ZTORQUE Molecule is out of bounds!");
	}
	return add_field(ctx, SMD_MOL_ZTORQUE, n_blocks, nullptr, c, 4, nullptr, 0, blocks, 3, "ZTORQUE Molecule");
}

extern "C" int smd_add_zpower(smd_ctx *ctx, int32_t n_blocks, const int32_t *blocks, const double c[2])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n_blocks >= 0 && (blocks || n_blocks == 0) && c, "bad ZPOWERPOTENTIAL arguments");
	for (int j = 0; j < n_blocks; j++) {
		long long start = blocks[2 * j], cnt = blocks[2 * j + 1];
		REQUIRE(start >= 0 && cnt >= 0 && start + cnt <= ctx->n_global, "ZPOWERPOTENTIAL Molecule is out of bounds!");
	}
	return add_field(ctx, SMD_MOL_ZPOWERPOTENTIAL, n_blocks, nullptr, c, 2, nullptr, 0, blocks, 2, "ZPOWERPOTENTIAL Molecule");
}

// ---- the kinds of MDsubstrate.cpp's switch (MDsubstrate.cpp:245-259): the caller decides which driver's switch it follows
extern "C" int smd_add_offset_boundary(smd_ctx *ctx, int32_t n, const int32_t *idx, const double c[4])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n >= 0 && (idx || n == 0) && c, "bad OFFSET_BOUNDARY arguments");
	int dim = (int)c[0];   // static_cast<int>(constants[0]), MD.h:512
	REQUIRE(dim >= 0 && dim <= 2, "OFFSET_BOUNDARY: constants[0] must name an axis (0, 1 or 2)");
	return add_field(ctx, SMD_MOL_OFFSET_BOUNDARY, n, idx, c, 4, nullptr, 0, nullptr, 0, "OFFSET_BOUNDARY Molecule");
}

extern "C" int smd_add_rigidbend(smd_ctx *ctx, int32_t n, const int32_t *ij, const double c[5])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n >= 0 && (ij || n == 0) && c, "bad RIGIDBEND arguments");
	if (ctx->slab) { ctx->err = "slab mode supports CHAIN, BOND and BEND molecules only"; return SMD_ERR_UNSUPPORTED; }
	int rc = check_index(ctx, ij, 2 * (size_t)n, "RIGIDBEND Molecule");
	if (rc) return rc;
	// records of two indices: uploaded as a flat list; the fifth constant travels in host_radius
	rc = add_field(ctx, SMD_MOL_RIGIDBEND, 2 * n, ij, c, 4, nullptr, 0, nullptr, 0, "RIGIDBEND Molecule");
	if (rc) return rc;
	ctx->fields.back().n = n;
	ctx->fields.back().host_radius.assign(1, c[4]);
	return SMD_OK;
}

extern "C" int smd_add_pullbead(smd_ctx *ctx, int32_t n, const int32_t *idx, const double c[4])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n >= 0 && (idx || n == 0) && c, "bad PULLBEAD arguments");
	return add_field(ctx, SMD_MOL_PULLBEAD, n, idx, c, 4, nullptr, 0, nullptr, 0, "PULLBEAD Molecule");
}

extern "C" int smd_add_nanocore(smd_ctx *ctx, int32_t n, const int32_t *idx, const double *C)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n >= 0 && (idx || n == 0) && (C || n == 0), "bad NANOCORE arguments");
	return add_field(ctx, SMD_MOL_NANOCORE, n, idx, nullptr, 0, C, 22 * (size_t)n, nullptr, 0, "NANOCORE Molecule");
}

// force (MODE 0) or potential (MODE 1; partial sums finished by `done(blocks, term)`) of one field molecule
template <int MODE, class Done>
static int launch_field(smd_ctx *ctx, const FieldMol &f, const Particle *pos, Done done)
{
	double *acc = MODE == 0 ? ctx->acc : nullptr, *part = MODE == 0 ? nullptr : ctx->partials;
	switch (f.kind) {
	case SMD_MOL_BOUNDARY:
		if (f.n <= 0) break;
		LAUNCHP(k_boundary<MODE>, nblk(f.n, TPB), TPB, 0, f.n, ctx->cap, pos, ctx->slot_of, ctx->geom, f.d_idx, (int)f.c[0], f.c[1], f.c[3], acc, part);
		done(nblk(f.n, TPB), SMD_TERM_FIELD);
		break;
	case SMD_MOL_FLOATING_BASE:
		if (f.n <= 0) break;
		LAUNCHP(k_floating_base<MODE>, nblk(f.n, TPB), TPB, 0, f.n, ctx->cap, pos, ctx->slot_of, f.d_idx, f.d_C, acc, part);
		done(nblk(f.n, TPB), SMD_TERM_FIELD);
		break;
	case SMD_MOL_ZTORQUE:
		for (int j = 0; j < f.n; j++) {
			int start = f.blocks[3 * j], nCh = f.blocks[3 * j + 1], len = f.blocks[3 * j + 2];
			long long nt = (long long)nCh * (len - 2);
			if (len < 3 || nt <= 0) continue;
			LAUNCHP(k_ztorque<MODE>, nblk((int)nt, TPB), TPB, 0, ctx->cap, pos, ctx->slot_of, ctx->geom, start, nCh, len, f.c[0], f.c[1], f.c[2], f.c[3],
			       acc, part);
			done(nblk((int)nt, TPB), SMD_TERM_FIELD);
		}
		break;
	case SMD_MOL_ZPOWERPOTENTIAL:
		for (int j = 0; j < f.n; j++) {
			int start = f.blocks[2 * j], cnt = f.blocks[2 * j + 1];
			if (cnt <= 0) continue;
			LAUNCHP(k_zpower<MODE>, nblk(cnt, TPB), TPB, 0, cnt, ctx->cap, pos, ctx->slot_of, start, f.c[0], f.c[1], acc, part);
			done(nblk(cnt, TPB), SMD_TERM_FIELD);
		}
		break;
	case SMD_MOL_OFFSET_BOUNDARY:
		if (f.n <= 0 || MODE != 0) break;
		LAUNCHP(k_offset_boundary, nblk(f.n, TPB), TPB, 0, f.n, ctx->cap, pos, ctx->slot_of, ctx->geom, f.d_idx, (int)f.c[0], f.c[1], f.c[2], f.c[3], acc);
		break;
	case SMD_MOL_RIGIDBEND:
		if (f.n <= 0 || MODE != 0) break;
		LAUNCHP(k_rigidbend, nblk(f.n, TPB), TPB, 0, f.n, ctx->cap, pos, ctx->slot_of, ctx->geom, f.d_idx, f.c[0], f.c[1], f.c[2], f.c[3],
		        f.host_radius[0], acc);
		break;
	case SMD_MOL_PULLBEAD:
		if (f.n <= 0 || MODE != 0) break;
		LAUNCHP(k_pullbead, nblk(ctx->N, TPB), TPB, 0, cnt_of(ctx), ctx->N, f.n, ctx->cap, pos, ctx->slot_of, ctx->geom, f.d_idx, f.c[0], f.c[1], f.c[2],
		        f.c[3], acc);
		break;
	default: break;
	}
	return SMD_OK;
}

// rebuild the assembled bead lists: molecule i sees its own beads followed by those of every later BEAD molecule
static int rebuild_bead_lists(smd_ctx *ctx)
{
	CK(cudaSetDevice(ctx->device));
	for (size_t i = 0; i < ctx->beads.size(); i++) {
		std::vector<int> all;
		for (size_t j = i; j < ctx->beads.size(); j++) all.insert(all.end(), ctx->beads[j].own.begin(), ctx->beads[j].own.end());
		BeadMol &b = ctx->beads[i];
		if (b.d_beads) cudaFree(b.d_beads);
		b.d_beads = nullptr;
		b.nAll = (int)all.size();
		if (!all.empty()) {
			CK(cudaMalloc(&b.d_beads, all.size() * sizeof(int)));
			CK(cudaMemcpy(b.d_beads, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice));
		}
	}
	return SMD_OK;
}

extern "C" int smd_add_beads(smd_ctx *ctx, int32_t n, const int32_t *idx, const double *C)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n >= 0 && (idx || n == 0) && C, "bad BEAD arguments");
	if (ctx->slab) { ctx->err = "slab mode supports CHAIN, BOND and BEND molecules only"; return SMD_ERR_UNSUPPORTED; }
	int rc = check_index(ctx, idx, (size_t)n, "BEAD Molecule");
	if (rc) return rc;
	CK(cudaSetDevice(ctx->device));
	BeadMol b;
	b.nOwn = n; b.nAll = n; b.d_beads = nullptr; b.d_C = nullptr;
	b.radius = C[4];   // BEADRADIUS, MD.h:53
	b.mol_index = ctx->n_molecules;
	b.own.assign(idx, idx + n);
	size_t nc = 22 * (size_t)ctx->nT * ctx->nT;
	CK(cudaMalloc(&b.d_C, nc * sizeof(double)));
	CK(cudaMemcpy(b.d_C, C, nc * sizeof(double), cudaMemcpyHostToDevice));
	ctx->beads.push_back(b);
	ctx->mol_order.push_back({SMD_MOL_BEAD, (int)ctx->beads.size() - 1, 1});
	ctx->n_molecules++;
	return rebuild_bead_lists(ctx);
}

extern "C" int smd_set_temperature(smd_ctx *ctx, double temperature)
{
	if (!ctx) return SMD_ERR_ARG;
	ctx->temperature = temperature;
	return SMD_OK;
}

extern "C" int smd_set_noise(smd_ctx *ctx, const double *u)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(u, "null noise");
	REQUIRE(ctx->desc.noise == SMD_NOISE_EXTERNAL, "context was not created with SMD_NOISE_EXTERNAL");
	CK(cudaSetDevice(ctx->device));
	// the previous Langevin kernel may still be reading the buffer: stream order takes care of it
	CK(cudaMemcpyAsync(ctx->noise, u, 3 * (size_t)ctx->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));   // host buffer may be reused by the caller
	ctx->noise_ready = true;
	return SMD_OK;
}

// ------------------------------------------------------------------------------------------------ cell build
// the window the CURRENT sorted order / start[] refer to, and the one the next tagging pass bins into
static inline int *cur_win(smd_ctx *ctx) { return ctx->win + ctx->wcur * WIN_WORDS; }
static inline int *next_win(smd_ctx *ctx) { return ctx->win + (ctx->wcur ^ 1) * WIN_WORDS; }

// what a kernel that moves the particles needs to fill the histogram of the following build itself (bin_particle); not
// in slab mode (received particles arrive after the pass) and not before a build has published the next window.
// The caller is about to launch such a pass: the histogram is pending from here on.
static BinArgs bin_args(smd_ctx *ctx)
{
	BinArgs b = {nullptr, nullptr, nullptr};
	if (ctx->hist_pending) drop_pending_histogram(ctx);   // a second pass without a build in between: that histogram is stale
	if (!ctx->prebin || (ctx->slab && ctx->no_slab_prebin) || !ctx->next_win_valid || !ctx->cells_valid) return b;
	b.win = next_win(ctx); b.count = ctx->count; b.cellOfSlot = ctx->cellOfSlot;
	ctx->hist_pending = true;
	return b;
}

// the particles are about to be replaced / rescaled / re-tagged by a pass that does not bin: drop a pending histogram
static int drop_pending_histogram(smd_ctx *ctx)
{
	if (ctx->hist_pending) LAUNCH(k_clear_count, 296, 256, 0, next_win(ctx), ctx->count);
	ctx->hist_pending = false;
	ctx->next_win_valid = false;
	return SMD_OK;
}

static int build_cells(smd_ctx *ctx)
{
	int N = ctx->N, cur = ctx->cur, nxt = cur ^ 1, pcur = ctx->pcur, pnxt = pcur ^ 1;
	// particles were tagged with their cell (and bbox[] accumulated) by whoever moved them last
	if (ctx->slab && !ctx->ext_valid)   // no unpack since the last build: the extended count is the current one
		CK(cudaMemcpyAsync(ctx->dN + 1, ctx->dN, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
	ctx->ext_valid = false;
	const bool prebinned = ctx->hist_pending;   // the kernel that moved the particles has filled the histogram (bin_particle)
	ctx->hist_pending = false;
	if (prebinned) ctx->wcur ^= 1;              // ... under the window the last build published for it: this build's window
	else {
		ProfScope ps(ctx, SMD_PHASE_BUILD_HIST);
		LAUNCHP(k_bin, nblk(N, TPB), TPB, 0, cnt_ext(ctx), ctx->pos[pcur], ctx->geom, ctx->bbox, ctx->cellcap, ctx->count, ctx->cellOfSlot, ctx->errflag,
		       ctx->gid[cur], ctx->slab ? ctx->slot_of : nullptr);
	}
	{
		ProfScope ps(ctx, SMD_PHASE_BUILD_SCAN);
		LAUNCHP(k_scan, SCAN_BLOCKS, SCAN_TPB, 0, ctx->count, ctx->bbox, ctx->geom, ctx->cellcap, cur_win(ctx), next_win(ctx),
		       prebinned ? 1 : 0, ctx->start, ctx->cursor, N, ctx->slab ? ctx->dN : nullptr, ctx->errflag, ctx->scan_state, (unsigned)(ctx->rebuilds + 1),
		       (const Particle *)ctx->pos[pcur], ctx->cellOfSlot, ctx->scan_barrier);
		ctx->next_win_valid = true;
	}
	{
		ProfScope ps(ctx, SMD_PHASE_BUILD_PLACE);
		LAUNCHP(k_place, nblk(N, TPB), TPB, 0, cnt_ext(ctx), ctx->cellOfSlot, ctx->cursor, ctx->order, ctx->gid[cur],
		        (ctx->slab && prebinned) ? ctx->slot_of : (int *)nullptr);
	}
	{
		ProfScope ps(ctx, SMD_PHASE_BUILD_REORDER);
		LAUNCHP(k_reorder, nblk(N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, ctx->order, ctx->cellOfSlot, ctx->start, ctx->pos[pcur], ctx->pos[pnxt],
		       ctx->vel[cur], ctx->vel[nxt], ctx->unw[cur], ctx->unw[nxt], ctx->acc_live ? ctx->acc : nullptr, ctx->acc2, ctx->gid[cur],
		       ctx->gid[nxt], ctx->slot_of, ctx->pos32, ctx->acut, ctx->bbox, (ctx->rebuilds & 255) == 255 ? 1 : 0, ctx->pos16, ctx->arad, cur_win(ctx),
		       ctx->geom);
	}
	if (ctx->acc_live) std::swap(ctx->acc, ctx->acc2);   // a build in between force evaluation and the next kick keeps acc aligned
	ctx->cur = nxt;
	ctx->pcur = pnxt;
	ctx->cells_valid = true;
	ctx->rebuilds++;
	return SMD_OK;
}


extern "C" int smd_build_cells(smd_ctx *ctx)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(ctx->particles_set, "smd_set_particles first");
	CK(cudaSetDevice(ctx->device));
	if (ctx->cells_valid) return SMD_OK;
	return build_cells(ctx);
}

// ------------------------------------------------------------------------------------------------ force terms
static int pair_smem(smd_ctx *ctx) { return 6 * ctx->nT * ctx->nT * (int)sizeof(double); }

// NOTE on acc[]: a build permutes pos / vel / gid but not acc (it is zeroed or recomputed right after every
// build in the reference's loop order), so forces are always evaluated as: [build] -> zero -> terms.
// blocks per bead of k_bead: a few beads are spread over the device row by row, many beads take a block each
static int bead_parts(int nbeads) { return std::max(1, std::min(32, 296 / std::max(nbeads, 1))); }

static int add_molecule_forces(smd_ctx *ctx, uint32_t mask)
{
	const Particle *pos = ctx->pos[ctx->pcur];
	if (mask & SMD_MASK(SMD_TERM_CHAIN))
		for (auto &cb : ctx->chains) {
			if (cb.nChains <= 0) continue;
			if (ctx->slab)
				LAUNCH(k_chain_slab<0>, nblk(ctx->N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, pos, ctx->gid[ctx->cur], ctx->slot_of, ctx->geom, cb,
				       ctx->acc, nullptr, 1.0, 1.0, 1.0, ctx->errflag);
			else
				LAUNCH(k_chain<0>, nblk(cb.nChains, TPB), TPB, 0, ctx->cap, pos, ctx->slot_of, ctx->geom, cb, ctx->acc, nullptr, 1.0, 1.0, 1.0);
		}
	if (mask & SMD_MASK(SMD_TERM_BOND))
		for (auto &b : ctx->bonds)
			if (b.n > 0)
				LAUNCHP(k_bond<0>, nblk(b.n, TPB), TPB, 0, b.n, ctx->cap, pos, ctx->slot_of, ctx->geom, b.d_ij, b.c[0], b.c[1], ctx->acc, nullptr, 1.0, 1.0, 1.0, ctx->gid[ctx->cur], ctx->errflag);
	if (mask & SMD_MASK(SMD_TERM_BEND))
		for (auto &b : ctx->bends)
			if (b.n > 0)
				LAUNCHP(k_bend<0>, nblk(b.n, TPB), TPB, 0, b.n, ctx->cap, pos, ctx->slot_of, ctx->geom, b.d_ijk, b.c[0], b.c[1], ctx->acc, nullptr, 1.0, 1.0, 1.0, ctx->gid[ctx->cur], ctx->errflag);
	if (mask & SMD_MASK(SMD_TERM_BALL))
		for (auto &b : ctx->balls)
			if (b.n > 0)
				LAUNCHP(k_ball<0>, nblk(b.n, TPB), TPB, 0, b.n, ctx->cap, pos, ctx->slot_of, ctx->geom, b.d_cj, b.c[0], b.c[1], ctx->acc, nullptr, 1.0, 1.0, 1.0);
	if (mask & SMD_MASK(SMD_TERM_BEAD))
		for (auto &b : ctx->beads) {
			if (b.nOwn <= 0) continue;
			LAUNCHP(k_beadbead<0>, 1, TPB, 0, b.nOwn, b.nAll, ctx->cap, pos, ctx->slot_of, ctx->geom, ctx->nT, b.d_beads, b.d_C, b.radius,
			       ctx->acc, nullptr, 1.0, 1.0, 1.0);
			LAUNCHP(k_bead<0>, b.nOwn * bead_parts(b.nOwn), TPB, 0, b.nOwn, ctx->cap, pos, ctx->gid[ctx->cur], ctx->slot_of, ctx->start, cur_win(ctx), ctx->geom, ctx->nT,
			       b.d_beads, b.d_C, b.nOwn <= 20 ? 1 : 0, 0, ctx->acc, nullptr, 1.0, 1.0, 1.0, bead_parts(b.nOwn));
		}
	if (mask & SMD_MASK(SMD_TERM_FIELD))
		for (auto &f : ctx->fields) {
			int rc = launch_field<0>(ctx, f, pos, [](int, int) {});
			if (rc) return rc;
		}
	if (mask & SMD_MASK(SMD_TERM_NANOCORE))
		for (auto &f : ctx->fields)
			if (f.kind == SMD_MOL_NANOCORE && f.n > 0)
				LAUNCHP(k_bead<0>, f.n * bead_parts(f.n), TPB, 0, f.n, ctx->cap, pos, ctx->gid[ctx->cur], ctx->slot_of, ctx->start, cur_win(ctx), ctx->geom, ctx->nT,
				       f.d_idx, f.d_C, 0, 1, ctx->acc, nullptr, 1.0, 1.0, 1.0, bead_parts(f.n));
	return SMD_OK;
}

static size_t field_sum_slots(const smd_ctx *ctx)
{
	size_t k = 0;
	for (auto &f : ctx->fields) k += (f.kind == SMD_MOL_ZTORQUE || f.kind == SMD_MOL_ZPOWERPOTENTIAL) ? (size_t)f.n : 1;
	return k;
}

static int bead_mass_divide(smd_ctx *ctx)
{
	for (auto &b : ctx->beads) {
		if (b.nOwn <= 0) continue;
		double mass = (4.0) * M_PI * b.radius * b.radius;   // MD.cpp:346
		LAUNCH(k_bead_mass, nblk(b.nOwn, 64), 64, 0, b.nOwn, ctx->cap, b.d_beads, ctx->slot_of, mass, ctx->acc);
	}
	return SMD_OK;
}

// NANOCORE beads are divided by their mass on restart (MD.cpp:290-303) and before the second half kick (:495-508), but
// not at the top of the step (:340-355 handles BEAD only)
static int nanocore_mass_divide(smd_ctx *ctx)
{
	for (auto &f : ctx->fields)
		if (f.kind == SMD_MOL_NANOCORE && f.n > 0)
			LAUNCH(k_nanocore_mass, nblk(f.n, 64), 64, 0, f.n, ctx->cap, f.d_idx, ctx->slot_of, f.d_C, ctx->acc);
	return SMD_OK;
}

// langevin.h:235; with per-type friction the reference computes sT[] ONCE, at the temperature of its first call
// (langevin.h:236-241: `if(gT!=NULL && sT==NULL)`), so a temperature ramp never reaches the noise amplitude
static double langevin_sigma(smd_ctx *ctx)
{
	if (ctx->sigma_frozen) {
		if (!(ctx->sigma_temperature >= 0)) ctx->sigma_temperature = ctx->temperature;   // first evaluation
		return sqrt((6.0 * ctx->sigma_temperature * ctx->desc.gamma) / ctx->desc.dt);
	}
	return sqrt((6.0 * ctx->temperature * ctx->desc.gamma) / ctx->desc.dt);
}

extern "C" int smd_set_gamma_type(smd_ctx *ctx, int32_t n_types, const double *gamma_type)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(gamma_type && n_types == ctx->nT, "gammaType needs one value per particle type");
	REQUIRE(gamma_type[0] > 0, "gammaType[0] must be positive");
	// Langevin::compute with gT != NULL (langevin.h:248-281): `int type=0;//p[i].type;` -- every particle gets gT[0], sT[0]
	ctx->desc.gamma = gamma_type[0];
	ctx->sigma_frozen = true;
	ctx->sigma_temperature = -1.0;
	return SMD_OK;
}

static int add_langevin(smd_ctx *ctx, int64_t step)
{
	double sigma = langevin_sigma(ctx);
	const double *ext = nullptr;
	if (ctx->desc.noise == SMD_NOISE_EXTERNAL) {
		REQUIRE(ctx->noise_ready, "SMD_NOISE_EXTERNAL: call smd_set_noise before every Langevin evaluation");
		ext = ctx->noise;
		ctx->noise_ready = false;
	}
	LAUNCH(k_langevin, nblk(ctx->N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, ctx->vel[ctx->cur], ctx->acc, ctx->gid[ctx->cur], ctx->desc.gamma, sigma,
	       ctx->desc.seed, (uint64_t)step, ext);
	return SMD_OK;
}

static int ready(smd_ctx *ctx)
{
	REQUIRE(ctx->particles_set, "smd_set_particles first");
	REQUIRE(ctx->tables_set, "smd_set_pair_tables first");
	CK(cudaSetDevice(ctx->device));
	return SMD_OK;
}

template <int SPLIT>
static int pair_force_smem_t(smd_ctx *ctx, bool du)   // du: the force + dPotential instance (EMODE 3) stages a second table
{
	typedef PairCfg<SPLIT> Cfg;
	return (int)((sizeof(PairSmemT<Cfg::BT>) + 15) & ~size_t(15)) + PTAB_STRIDE * ctx->nT * ctx->nT * (int)sizeof(double) +
	       (du ? (PTAB_STRIDE * ctx->nT * ctx->nT + Cfg::BT) * (int)sizeof(double) : 0) +
	       (Cfg::BT / 32) * Cfg::CAP * 32 * (int)sizeof(unsigned short);
}
static int pair_force_smem(smd_ctx *ctx, bool du) { return pair_force_smem_t<1>(ctx, du); }

// Three threads per particle (k_pair_force2<.., SPLIT = 3>): a third of the critical path per thread.  Worth it where the grid
// of the one-thread engine does not fill the device, i.e. where the launch takes as long as its slowest thread
// (profiles/r02b_pair3_ab.md, profiles/r02d_timeline.md).
static int pair_split(const smd_ctx *ctx)
{
	if (!ctx->tables_symmetric) return 1;
	if (ctx->pair3 >= 0) return ctx->pair3 ? 3 : 1;   // SMD_PAIR3 = 0 | 1
	return ctx->N <= ctx->pair3_max ? 3 : 1;
}
// Which blocks of particles go to which engine.  Small systems: everything to the three-thread engine.  Large ones: the
// one-thread engine -- and, as an experiment (SMD_PAIR_TAIL=1, see launch_pair_force), in whole rounds of (resident blocks
// per SM x SMs) blocks with the particles of a last round that would be less than half full handed to the three-thread engine
// in a second launch behind the first.
struct PairPlan { int nb_main, first_tail, nb_tail; };
static PairPlan pair_plan(const smd_ctx *ctx)
{
	const int N = ctx->N;
	PairPlan p = {nblk(N, PAIR_TPB), 0, 0};
	const int split = pair_split(ctx);
	if (split == 3) { p.nb_main = 0; p.first_tail = 0; p.nb_tail = nblk(N, PAIR_SPLIT_NP); return p; }
	if (!ctx->tables_symmetric || !ctx->pair_tail || ctx->slab) return p;
	const int slots = ctx->sm_count * SMD_PAIR_BLOCKS;
	const int rem = p.nb_main % slots;
	if (p.nb_main > slots && rem > 0 && 2 * rem < slots) {
		p.nb_main -= rem;
		p.first_tail = p.nb_main * PAIR_TPB;
		p.nb_tail = nblk(N - p.first_tail, PAIR_SPLIT_NP);
	}
	return p;
}

template <int EMODE, bool LANGEVIN, int SPLIT>
static int launch_pair_split(smd_ctx *ctx, const LangevinArgs &lg, const EnergyArgs &en, int first, int nb, int part0, int nowait = 0)
{
	typedef PairCfg<SPLIT> Cfg;
	// > 48 KB of dynamic shared memory (many particle types) needs the opt-in: a property of the function ON A DEVICE, raised once
	static bool attr_done[64] = {false};
	const int smem = pair_force_smem_t<SPLIT>(ctx, EMODE == 3);
	if (smem > 48 * 1024 && !attr_done[ctx->device & 63]) {
		CK(cudaFuncSetAttribute(k_pair_force2<EMODE, LANGEVIN, true, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
		attr_done[ctx->device & 63] = true;
	}
	PairGeo pg = ctx->pgeo;
	pg.first = first; pg.part0 = part0; pg.nowait = nowait;
	LAUNCHP((k_pair_force2<EMODE, LANGEVIN, true, SPLIT>), nb, Cfg::BT, smem, cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->pos32,
	       ctx->start, cur_win(ctx), ctx->geom, ctx->nT, ctx->fC, ctx->ptab, pg, ctx->acc, lg, ctx->gid[ctx->cur], en, ctx->pos16);
	return SMD_OK;
}

// the pair force of all particles into acc[] (LANGEVIN: a = thermostat term + pair sum, else a += pair sum)
template <bool LANGEVIN>
static int launch_pair_force(smd_ctx *ctx, const LangevinArgs &lg)
{
	const PairPlan pl = pair_plan(ctx);
	PairGeo pg = ctx->pgeo;
	pg.first = 0; pg.part0 = 0;
	int rc = SMD_OK;
	// Hybrid (EXPERIMENT, SMD_PAIR_TAIL=1): the main launch in whole rounds, then a small three-thread launch for what would have
	// been the partial last round.  The small launch is a programmatic dependent of the main one, whose results it does not
	// need -- it needs the cell build, which had completed and flushed before the first main block got past its
	// griddepcontrol.wait, and no small block is resident before every main block has started -- so it does not wait
	// (PairGeo::nowait) and its blocks move in as main blocks leave.  Measured on C2: the pair kernels end at 178.0 instead of
	// 184.0 us of the step, 203.6 -> 200.4 us per MD step (the other order, small launch first: 201.5).  Off by default: 1.6 % do
	// not pay for a grid that reads the build's output without having waited on anything -- the read-only data path may keep
	// lines of the previous step that only a completed dependency is documented to invalidate.
	const int hybrid_nowait = (pl.nb_tail > 0 && pl.nb_main > 0 && pg.done != nullptr) ? 1 : 0;
	if (LANGEVIN && ctx->du_armed && ctx->tables_symmetric) {   // forces + Langevin + the dPotential of the box move proposed for this configuration (smd_step_mc)
		ctx->du_armed = false;
		if (pl.nb_main > 0)
			LAUNCHP((k_pair_force2<3, true, true>), pl.nb_main, PAIR_TPB, pair_force_smem(ctx, true), cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->pos32,
			       ctx->start, cur_win(ctx), ctx->geom, ctx->nT, ctx->fC, ctx->ptab, pg, ctx->acc, lg, ctx->gid[ctx->cur], ctx->du_en, ctx->pos16);
		if (pl.nb_tail > 0 && (rc = launch_pair_split<3, true, 3>(ctx, lg, ctx->du_en, pl.first_tail, pl.nb_tail, pl.nb_main, hybrid_nowait))) return rc;
		ctx->du_nparts = pl.nb_main + pl.nb_tail;
		ctx->du_ready = true;
		return SMD_OK;
	}
	if (!ctx->tables_symmetric) {
		LAUNCH((k_pair_force2<0, LANGEVIN, false>), pl.nb_main, PAIR_TPB, pair_force_smem(ctx, false), cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->pos32,
		       ctx->start, cur_win(ctx), ctx->geom, ctx->nT, ctx->fC, ctx->ptab, pg, ctx->acc, lg, ctx->gid[ctx->cur], EnergyArgs{}, ctx->pos16);
		return SMD_OK;
	}
	if (pl.nb_main > 0)
		LAUNCHP((k_pair_force2<0, LANGEVIN, true>), pl.nb_main, PAIR_TPB, pair_force_smem(ctx, false), cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->pos32,
		       ctx->start, cur_win(ctx), ctx->geom, ctx->nT, ctx->fC, ctx->ptab, pg, ctx->acc, lg, ctx->gid[ctx->cur], EnergyArgs{}, ctx->pos16);
	if (pl.nb_tail > 0) rc = launch_pair_split<0, LANGEVIN, 3>(ctx, lg, EnergyArgs{}, pl.first_tail, pl.nb_tail, pl.nb_main, hybrid_nowait);
	return rc;
}

static int forces(smd_ctx *ctx, uint32_t mask, int64_t step, bool langevin_first)
{
	int N = ctx->N;
	// CellOpt::build: always rebuilt (the reference rebuilds every step, MD.cpp:412)
	ctx->acc_live = false;   // about to be overwritten: no need to carry it through the build
	if (ctx->slab && ctx->exch_pending) {
		int rcx = smd_slab_exchange_recv(ctx);
		if (rcx) return rcx;
	}
	if (!ctx->cells_valid) { ProfScope ps(ctx, SMD_PHASE_BUILD); build_cells(ctx); }
	int rc;
	LangevinArgs lg = {};
	bool pair = mask & SMD_MASK(SMD_TERM_PAIR), lang = mask & SMD_MASK_LANGEVIN;
	if (langevin_first && lang && pair) {
		// steady-state step (MD.cpp:357-413): a = 0, thermostat, pair force in ONE kernel
		if (ctx->desc.noise == SMD_NOISE_EXTERNAL) {
			REQUIRE(ctx->noise_ready, "SMD_NOISE_EXTERNAL: call smd_set_noise before every Langevin evaluation");
			lg.ext_noise = ctx->noise;
			ctx->noise_ready = false;
		}
		lg.gamma = ctx->desc.gamma;
		lg.sigma = langevin_sigma(ctx);
		lg.seed = ctx->desc.seed; lg.step = (uint64_t)step;
		lg.vel = ctx->vel[ctx->cur]; lg.gid = ctx->gid[ctx->cur];
		// (a launch that also sums a dPotential is timed apart: bench.py's roofline describes the plain force launch; when
		// only the plain phase is being profiled, the other one inherits the request)
		const bool du_launch = ctx->du_armed && ctx->tables_symmetric;
		if (du_launch && ((ctx->prof_mask >> SMD_PHASE_PAIR) & 1u)) ctx->prof_mask |= 1u << SMD_PHASE_PAIR_DU;
		ProfScope ps(ctx, du_launch ? SMD_PHASE_PAIR_DU : SMD_PHASE_PAIR);
		if ((rc = launch_pair_force<true>(ctx, lg))) return rc;
		ctx->acc_live = true;
	} else {
		LAUNCH(k_zero3, nblk(N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, ctx->acc);
		ctx->acc_live = true;
		if (langevin_first && lang) {
			ProfScope ps(ctx, SMD_PHASE_LANGEVIN);
			if ((rc = add_langevin(ctx, step))) return rc;
		}
		if (pair) {
			ProfScope ps(ctx, SMD_PHASE_PAIR);
			if ((rc = launch_pair_force<false>(ctx, lg))) return rc;
		}
		if (!langevin_first && lang) {
			ProfScope ps(ctx, SMD_PHASE_LANGEVIN);
			if ((rc = add_langevin(ctx, step))) return rc;
		}
	}
	ProfScope ps(ctx, SMD_PHASE_MOLECULES);
	return add_molecule_forces(ctx, mask);
}

extern "C" int smd_compute_forces(smd_ctx *ctx, uint32_t term_mask, int64_t step)
{
	if (!ctx) return SMD_ERR_ARG;
	int rc = ready(ctx);
	if (rc) return rc;
	// MD.cpp:192-262: pair, thermostat, molecules
	rc = forces(ctx, term_mask, step, false);
	return rc;
}

extern "C" int smd_resume(smd_ctx *ctx)
{
	if (!ctx) return SMD_ERR_ARG;
	int rc = ready(ctx);
	if (rc) return rc;
	bead_mass_divide(ctx);
	nanocore_mass_divide(ctx);
	LAUNCH(k_verlet_second, nblk(ctx->N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->vel[ctx->cur], ctx->acc, ctx->desc.dt);
	return SMD_OK;
}

extern "C" int smd_step_begin(smd_ctx *ctx, int64_t step)
{
	if (!ctx) return SMD_ERR_ARG;
	ctx->du_ready = false;   // the particles change: block sums armed by smd_arm_dpotential no longer describe them
	(void)step;
	int rc = ready(ctx);
	if (rc) return rc;
	int N = ctx->N;
	ProfScope ps(ctx, SMD_PHASE_INTEGRATE1);
	bead_mass_divide(ctx);                                                         // MD.cpp:340-355
	LAUNCH(k_verlet_first, nblk(N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->vel[ctx->cur], ctx->acc, ctx->unw[ctx->cur], ctx->geom,
	       ctx->desc.dt, ctx->bbox, ctx->errflag, ctx->gid[ctx->cur], bin_args(ctx));             // MD.cpp:356
	// a = 0 (MD.cpp:357-366) is folded into the force evaluation that follows: it overwrites a[]
	ctx->acc_live = false;
	ctx->cells_valid = false;
	if (ctx->slab) return smd_slab_exchange_send(ctx);   // migrants + halo on their way while the host enqueues the rest
	return SMD_OK;
}

extern "C" int smd_step_end(smd_ctx *ctx, int64_t step)
{
	if (!ctx) return SMD_ERR_ARG;
	int rc = ready(ctx);
	if (rc) return rc;
	// MD.cpp:410-478: thermostat, build, pair, molecules
	rc = forces(ctx, SMD_MASK_ALL, step, true);
	if (rc) return rc;
	{ ProfScope ps(ctx, SMD_PHASE_MOLECULES); bead_mass_divide(ctx); nanocore_mass_divide(ctx); }   // MD.cpp:480-508
	ProfScope ps(ctx, SMD_PHASE_INTEGRATE2);
	LAUNCH(k_verlet_second, nblk(ctx->N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->vel[ctx->cur], ctx->acc, ctx->desc.dt);   // :511
	return SMD_OK;
}

// CHAIN-only systems: the seam between two consecutive steps (chain forces, Verlet::second, next Verlet::first) is
// one kernel, see k_chain_kick
// The fused step seam (k_chain_kick) gathers the CHAIN terms per particle and applies both half kicks; every other
// molecule kind (BOND, BEND, BALL, BEAD, NANOCORE, the one-body fields) scatters into a[] with atomics BEFORE it.  The
// continuum-sphere particles of BEAD / NANOCORE molecules are divided by their mass inside the seam (BeadSet).
static bool bead_set(const smd_ctx *ctx, BeadSet &bs)
{
	bs.n = 0;
	auto add = [&](int id, double R, int twice) {
		for (int k = 0; k < bs.n; k++) if (bs.id[k] == id) return false;   // listed twice: divided twice, leave it to the plain path
		if (bs.n >= MAX_FUSED_BEADS) return false;
		bs.id[bs.n] = id; bs.twice[bs.n] = twice; bs.mass[bs.n] = (4.0) * M_PI * R * R;   // MD.cpp:346
		bs.n++;
		return true;
	};
	for (auto &b : ctx->beads)
		for (int id : b.own) if (!add(id, b.radius, 1)) return false;
	for (auto &f : ctx->fields)
		if (f.kind == SMD_MOL_NANOCORE)
			for (int j = 0; j < f.n; j++) if (!add(f.host_idx[j], f.host_radius[j], 0)) return false;
	return true;
}
static bool can_fuse(const smd_ctx *ctx)
{
	BeadSet bs;
	return !ctx->no_fuse && bead_set(ctx, bs) && (int)ctx->chains.size() <= MAX_FUSED_CHAINS && ctx->desc.noise != SMD_NOISE_EXTERNAL;
}
static bool only_chains(const smd_ctx *ctx)
{
	return ctx->bonds.empty() && ctx->bends.empty() && ctx->balls.empty() && ctx->fields.empty() && ctx->beads.empty();
}
// launch with the programmatic-stream-serialization attribute: the kernel may become resident before its predecessor in the
// stream has completed, and orders itself against it on the device (k_chain_kick: per-block completion words)
template <class... KArgs, class... Args>
static cudaError_t launch_dependent(void (*kernel)(KArgs...), int grid, int block, cudaStream_t stream, Args... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static ChainSet chain_set(const smd_ctx *ctx)
{
	ChainSet cs;
	cs.n = 0;
	for (auto &cb : ctx->chains)
		if (cb.nChains > 0) cs.b[cs.n++] = cb;
	return cs;
}

extern "C" int smd_step(smd_ctx *ctx, int64_t first_step, int32_t nsteps)
{
	if (!ctx) return SMD_ERR_ARG;
	NvtxRange nv("smd_step");
	REQUIRE(ctx->desc.noise != SMD_NOISE_EXTERNAL || nsteps <= 1, "external noise: one step per smd_set_noise");
	if (nsteps <= 0) return SMD_OK;
	struct Disarm { smd_ctx *c; ~Disarm() { c->du_for_last = false; c->du_armed = false; } } disarm{ctx};   // one call only
	ctx->du_ready = false;   // the particles are about to move
	if (!can_fuse(ctx)) {
		for (int k = 0; k < nsteps; k++) {
			ProfScope ps(ctx, SMD_PHASE_STEP);
			int rc = smd_step_begin(ctx, first_step + k);
			if (rc) return rc;
			rc = smd_step_end(ctx, first_step + k);
			if (rc) return rc;
		}
		return SMD_OK;
	}
	int rc = smd_step_begin(ctx, first_step);
	if (rc) return rc;
	const ChainSet cs = chain_set(ctx);
	BeadSet bs;
	bead_set(ctx, bs);
	const int N = ctx->N;
	// molecule kinds that add to a[] before the seam (none for CHAIN-only systems)
	const uint32_t scatter = only_chains(ctx) ? 0u : (SMD_MASK_ALL_MOLECULES & ~SMD_MASK(SMD_TERM_CHAIN));
	static_assert(TPB == PAIR_TPB, "seam block b waits for the completion word of pair block b: same slots");
	for (int k = 0; k < nsteps; k++) {
		ProfScope ps(ctx, SMD_PHASE_STEP);
		const bool last = (k == nsteps - 1);
		if (ctx->timeline) {   // smd_timeline: stamps of step nsteps - 2 (build, pair and a seam that is not the last of the call)
			static const int tl_on = 1, tl_off = 0;   // (stream-ordered copies into constant memory: the kernels launched after them see the new value)
			if (k == nsteps - 2) {
				k_tl_reset<<<1, 1, 0, ctx->stream>>>();
				CK(cudaMemcpyToSymbolAsync(c_tl_on, &tl_on, sizeof(int), 0, cudaMemcpyHostToDevice, ctx->stream));
			}
			if (last) CK(cudaMemcpyToSymbolAsync(c_tl_on, &tl_off, sizeof(int), 0, cudaMemcpyHostToDevice, ctx->stream));
		}
		ctx->du_armed = last && ctx->du_for_last;
		// The seam as a programmatic dependent of the pair kernel (CHAIN-only systems, nothing recorded in between): its
		// blocks move in where the tail of the pair grid has left SMs empty and start on their 128 slots as soon as the
		// pair block of those slots has signalled (PairGeo::done) -- the seam hides in the pair kernel's last, sparse round.
		const bool pdl = ctx->pdl && scatter == 0 && ctx->prof_mask == 0 && ctx->tables_symmetric;
		if (pdl) {
			if (!ctx->pair_done) {
				CK(cudaMalloc(&ctx->pair_done, (size_t)nblk(ctx->cap, 32) * sizeof(int)));   // (one word per block of either pair engine)
				CK(cudaMemsetAsync(ctx->pair_done, 0, (size_t)nblk(ctx->cap, 32) * sizeof(int), ctx->stream));
			}
			ctx->pgeo.done = ctx->pair_done;
			ctx->pgeo.epoch = ++ctx->pair_epoch;
		}
		// MD.cpp:410-413: thermostat, build, pair force (one kernel); then the scattering molecule kinds, if any
		rc = forces(ctx, SMD_MASK(SMD_TERM_PAIR) | SMD_MASK_LANGEVIN | scatter, first_step + k, true);
		ctx->pgeo.done = nullptr;
		if (rc) return rc;
		const int *done = pdl ? ctx->pair_done : nullptr;
		const int epoch = ctx->pair_epoch;
		const int done_per = TPB / PAIR_SPLIT_NP;   // completion words per seam block: one per 32 slots, whatever the pair engine
		ProfScope pf(ctx, SMD_PHASE_FUSED);
		if (last && pdl) {
			CK(launch_dependent(k_chain_kick<true>, nblk(N, TPB), TPB, ctx->stream, cnt_of(ctx), ctx->cap, (const Particle *)ctx->pos[ctx->pcur],
			                    (Particle *)nullptr, ctx->vel[ctx->cur], ctx->acc, ctx->unw[ctx->cur], (const int *)ctx->gid[ctx->cur],
			                    (const int *)ctx->slot_of, ctx->geom, cs, ctx->desc.dt, ctx->bbox, ctx->errflag, bs, 0, SlabComm{}, 0, (int *)nullptr,
			                    done, epoch, BinArgs{}, done_per));
			ctx->launches++;
		} else if (last) {
			LAUNCHP(k_chain_kick<true>, nblk(N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], (Particle *)nullptr, ctx->vel[ctx->cur],
			       ctx->acc, ctx->unw[ctx->cur], ctx->gid[ctx->cur], ctx->slot_of, ctx->geom, cs, ctx->desc.dt, ctx->bbox, ctx->errflag, bs, 0, SlabComm{}, 0, (int *)nullptr, (const int *)nullptr, 0, BinArgs{}, 1);
		} else {
			// slab mode: the seam kernel is also the send side of the exchange (migrants + halo packed as the particles get
			// their new positions, written straight into the neighbours' buffers); SMD_NO_SEAM_PACK=1: separate pack kernel
			const bool seam_pack = ctx->slab && !ctx->no_seam_pack;
			if (seam_pack) {
				REQUIRE(ctx->peer_set[0] && ctx->peer_set[1], "slab: connect both neighbours first (smd_slab_connect_*)");
				REQUIRE(!ctx->exch_pending, "slab: previous exchange not received yet");
				ctx->xseq++;
			}
			const BinArgs bin = bin_args(ctx);
			if (pdl) {
				CK(launch_dependent(k_chain_kick<false>, nblk(N, TPB), TPB, ctx->stream, cnt_of(ctx), ctx->cap, (const Particle *)ctx->pos[ctx->pcur],
				                    ctx->pos[ctx->pcur ^ 1], ctx->vel[ctx->cur], ctx->acc, ctx->unw[ctx->cur], (const int *)ctx->gid[ctx->cur],
				                    (const int *)ctx->slot_of, ctx->geom, cs, ctx->desc.dt, ctx->bbox, ctx->errflag, bs, 0,
				                    seam_pack ? ctx->comm : SlabComm{}, seam_pack ? ctx->xseq : 0, seam_pack ? ctx->gid[ctx->cur] : (int *)nullptr,
				                    done, epoch, bin, done_per));
				ctx->launches++;
			} else
			LAUNCHP(k_chain_kick<false>, nblk(N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->pos[ctx->pcur ^ 1],
			       ctx->vel[ctx->cur], ctx->acc, ctx->unw[ctx->cur], ctx->gid[ctx->cur], ctx->slot_of, ctx->geom, cs, ctx->desc.dt, ctx->bbox,
			       ctx->errflag, bs, 0, seam_pack ? ctx->comm : SlabComm{}, seam_pack ? ctx->xseq : 0, seam_pack ? ctx->gid[ctx->cur] : (int *)nullptr,
			       (const int *)nullptr, 0, bin, 1);
			ctx->pcur ^= 1;
			ctx->acc_live = false;
			ctx->cells_valid = false;
			if (seam_pack) ctx->exch_pending = true;
			else if (ctx->slab && (rc = smd_slab_exchange_send(ctx))) return rc;
		}
	}
	return SMD_OK;
}

// ------------------------------------------------------------------------------------------------ energies
static int finish_sum(smd_ctx *ctx, int nparts, int slot, double factor)
{
	LAUNCH(k_final_sum, 1, 256, 0, nparts, ctx->partials, ctx->scalars, slot, factor);
	return SMD_OK;
}

// MODE 1 potential, 2 dPotential; results accumulate on the host per term
template <int MODE>
static int energy_terms(smd_ctx *ctx, const double scale[3], double *out_terms, bool on_device = false)
{
	int rc = ready(ctx);
	if (rc) return rc;
	int N = ctx->N;
	if (ctx->slab && ctx->exch_pending) { ctx->err = "slab: energy call between smd_step_begin and smd_step_end"; return SMD_ERR_ARG; }
	if (!ctx->cells_valid) build_cells(ctx);   // dataExtraction::compute rebuilds its own CellOpt (dataExtraction.h:839-841)
	double sx = scale ? scale[0] : 1.0, sy = scale ? scale[1] : 1.0, sz = scale ? scale[2] : 1.0;
	const Particle *pos = ctx->pos[ctx->pcur];
	std::vector<int> term_of_slot;
	int slot = 0;
	auto push = [&](int term) { term_of_slot.push_back(term); return slot++; };
	REQUIRE(1 + ctx->chains.size() + ctx->bonds.size() + ctx->bends.size() + ctx->balls.size() + 2 * ctx->beads.size() + field_sum_slots(ctx) <= 64,
	        "too many molecule records for one energy call");
	if (MODE == 2 && ctx->du_ready && sx == ctx->du_en.sx && sy == ctx->du_en.sy && sz == ctx->du_en.sz) {
		// the block sums of this very dPotential were left behind by the force kernel of the last step (smd_arm_dpotential)
		ctx->du_ready = false;
		LAUNCH(k_final_sum, 1, 256, 0, ctx->du_nparts, ctx->du_partials, ctx->scalars, push(SMD_TERM_PAIR), 1.0);
	} else if (ctx->tables_symmetric && !ctx->force_onephase_energy) {
		// two-phase kernel, every unordered pair once (see k_pair_force2)
		int nb = nblk(N, PAIR_TPB);
		EnergyArgs en;
		en.sx = sx; en.sy = sy; en.sz = sz; en.partials = ctx->partials;
		double grow = 0;   // the most a component-wise scaling moves r^2 across a cutoff, relative
		for (double sc : {sx, sy, sz}) grow = std::max(grow, std::max(fabs(sc * sc - 1.0), fabs(1.0 / (sc * sc) - 1.0)));
		en.extra32 = MODE == 1 ? 0.0f : nextafterf((float)(1.01 * grow * ctx->geom.rc2 + 1e-7), INFINITY);
		LAUNCH((k_pair_force2<MODE, false, true>), nb, PAIR_TPB, pair_force_smem(ctx, false), cnt_of(ctx), ctx->cap, pos, ctx->pos32, ctx->start,
		       cur_win(ctx), ctx->geom, ctx->nT, ctx->uC, ctx->utab, ctx->pgeo, nullptr, LangevinArgs{}, ctx->gid[ctx->cur], en, ctx->pos16);
		finish_sum(ctx, nb, push(SMD_TERM_PAIR), 1.0);
	} else {
		int nb = nblk(N, TPB);
		LAUNCH(k_pair<(MODE == 1 ? PAIR_POTENTIAL : PAIR_DPOTENTIAL)>, nb, TPB, pair_smem(ctx), cnt_of(ctx), ctx->cap, pos, ctx->gid[ctx->cur], ctx->start,
		       cur_win(ctx), ctx->geom, ctx->nT, ctx->uC, nullptr, ctx->partials, nullptr, sx, sy, sz);
		finish_sum(ctx, nb, push(SMD_TERM_PAIR), 1.0);
	}
	for (auto &cb : ctx->chains) {
		if (cb.nChains <= 0) continue;
		int nb = ctx->slab ? nblk(N, TPB) : nblk(cb.nChains, TPB);
		if (ctx->slab)
			LAUNCH(k_chain_slab<MODE>, nb, TPB, 0, cnt_of(ctx), ctx->cap, pos, ctx->gid[ctx->cur], ctx->slot_of, ctx->geom, cb, nullptr, ctx->partials,
			       sx, sy, sz, ctx->errflag);
		else
			LAUNCH(k_chain<MODE>, nb, TPB, 0, ctx->cap, pos, ctx->slot_of, ctx->geom, cb, nullptr, ctx->partials, sx, sy, sz);
		finish_sum(ctx, nb, push(SMD_TERM_CHAIN), 1.0);
	}
	for (auto &b : ctx->bonds) {
		if (b.n <= 0) continue;
		int nb = nblk(b.n, TPB);
		LAUNCH(k_bond<MODE>, nb, TPB, 0, b.n, ctx->cap, pos, ctx->slot_of, ctx->geom, b.d_ij, b.c[0], b.c[1], nullptr, ctx->partials, sx, sy, sz, ctx->gid[ctx->cur], ctx->errflag);
		finish_sum(ctx, nb, push(SMD_TERM_BOND), 1.0);
	}
	for (auto &b : ctx->bends) {
		if (b.n <= 0) continue;
		int nb = nblk(b.n, TPB);
		LAUNCH(k_bend<MODE>, nb, TPB, 0, b.n, ctx->cap, pos, ctx->slot_of, ctx->geom, b.d_ijk, b.c[0], b.c[1], nullptr, ctx->partials, sx, sy, sz, ctx->gid[ctx->cur], ctx->errflag);
		finish_sum(ctx, nb, push(SMD_TERM_BEND), 1.0);
	}
	for (auto &b : ctx->balls) {
		if (b.n <= 0) continue;
		int nb = nblk(b.n, TPB);
		LAUNCH(k_ball<MODE>, nb, TPB, 0, b.n, ctx->cap, pos, ctx->slot_of, ctx->geom, b.d_cj, b.c[0], b.c[1], nullptr, ctx->partials, sx, sy, sz);
		finish_sum(ctx, nb, push(SMD_TERM_BALL), 1.0);
	}
	for (auto &b : ctx->beads) {
		if (b.nOwn <= 0) continue;
		LAUNCH(k_beadbead<MODE>, 1, TPB, 0, b.nOwn, b.nAll, ctx->cap, pos, ctx->slot_of, ctx->geom, ctx->nT, b.d_beads, b.d_C, b.radius, nullptr,
		       ctx->partials, sx, sy, sz);
		finish_sum(ctx, 1, push(SMD_TERM_BEAD), 1.0);
		LAUNCH(k_bead<MODE>, b.nOwn * bead_parts(b.nOwn), TPB, 0, b.nOwn, ctx->cap, pos, ctx->gid[ctx->cur], ctx->slot_of, ctx->start, cur_win(ctx), ctx->geom, ctx->nT, b.d_beads, b.d_C, 0, 0,
		       nullptr, ctx->partials, sx, sy, sz, bead_parts(b.nOwn));
		finish_sum(ctx, b.nOwn * bead_parts(b.nOwn), push(SMD_TERM_BEAD), 1.0);
	}
	for (auto &f : ctx->fields) {
		if (f.kind == SMD_MOL_NANOCORE) {
			if (f.n <= 0) continue;
			LAUNCH(k_bead<MODE>, f.n * bead_parts(f.n), TPB, 0, f.n, ctx->cap, pos, ctx->gid[ctx->cur], ctx->slot_of, ctx->start, cur_win(ctx), ctx->geom, ctx->nT, f.d_idx, f.d_C, 0, 1,
			       nullptr, ctx->partials, sx, sy, sz, bead_parts(f.n));
			finish_sum(ctx, f.n * bead_parts(f.n), push(SMD_TERM_NANOCORE), 1.0);
		} else if (MODE == 1) {   // the one-body fields have no dPotential (MD.cpp:642-669)
			int frc = launch_field<1>(ctx, f, pos, [&](int nb, int term) { finish_sum(ctx, nb, push(term), 1.0); });
			if (frc) return frc;
		}
	}
	if (on_device) {   // the terms stay on the device (ctx->terms_dev), nothing waits: smd_dpotential_device
		SlotTerms st;
		st.n = slot;
		for (int k = 0; k < slot; k++) st.term[k] = (signed char)term_of_slot[k];
		if (!ctx->terms_dev) CK(cudaMalloc(&ctx->terms_dev, SMD_NTERMS * sizeof(double)));
		LAUNCH(k_fold_terms, 1, 32, 0, st, ctx->scalars, ctx->terms_dev, SMD_NTERMS);
		return SMD_OK;
	}
	CK(cudaMemcpyAsync(ctx->h_pinned, ctx->scalars, slot * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	rc = check_device_errors(ctx);
	if (rc) return rc;
	for (int t = 0; t < SMD_NTERMS; t++) out_terms[t] = 0;
	for (int k = 0; k < slot; k++) out_terms[term_of_slot[k]] += ctx->h_pinned[k];
	return SMD_OK;
}

extern "C" int smd_potential(smd_ctx *ctx, double *out_terms)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(out_terms, "null output");
	return energy_terms<1>(ctx, nullptr, out_terms);
}

extern "C" int smd_dpotential(smd_ctx *ctx, const double scale[3], double *out_terms)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(out_terms && scale, "null argument");
	return energy_terms<2>(ctx, scale, out_terms);
}

extern "C" int smd_dpotential_device(smd_ctx *ctx, const double scale[3], double **d_terms)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(d_terms && scale, "null argument");
	NvtxRange nv("smd_dpotential_device");
	int rc = energy_terms<2>(ctx, scale, nullptr, true);
	*d_terms = ctx->terms_dev;
	return rc;
}

extern "C" int smd_kinetic(smd_ctx *ctx, double *out)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(out && ctx->particles_set, "bad call");
	CK(cudaSetDevice(ctx->device));
	int nb = std::min(nblk(ctx->N, 256), 1024);
	LAUNCH(k_kinetic, nb, 256, 0, cnt_of(ctx), ctx->cap, ctx->vel[ctx->cur], ctx->gid[ctx->cur], ctx->partials);
	finish_sum(ctx, nb, 0, 1.0);
	CK(cudaMemcpyAsync(ctx->h_pinned, ctx->scalars, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	int rc = check_device_errors(ctx);
	if (rc) return rc;
	*out = ctx->h_pinned[0];
	return SMD_OK;
}

extern "C" int smd_count_pairs(smd_ctx *ctx, int64_t *total, int32_t *per_particle)
{
	if (!ctx) return SMD_ERR_ARG;
	int rc = ready(ctx);
	if (rc) return rc;
	int N = ctx->N;
	if (!ctx->cells_valid) build_cells(ctx);
	int nb = nblk(N, TPB);
	if (ctx->slab && per_particle) { ctx->err = "slab: per-particle counts are not exported (use smd_slab_get_local)"; return SMD_ERR_UNSUPPORTED; }
	if (ctx->slab) CK(cudaMemsetAsync(ctx->icount, 0, (size_t)ctx->cap * sizeof(int), ctx->stream));
	LAUNCH(k_pair<PAIR_COUNT>, nb, TPB, pair_smem(ctx), cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->gid[ctx->cur], ctx->start, cur_win(ctx), ctx->geom,
	       ctx->nT, ctx->fC, nullptr, ctx->partials, ctx->icount, 1.0, 1.0, 1.0);
	finish_sum(ctx, nb, 0, 1.0);
	CK(cudaMemcpyAsync(ctx->h_pinned, ctx->scalars, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	if (per_particle) {
		LAUNCH(k_export_int, nblk(N, TPB), TPB, 0, N, ctx->icount, ctx->gid[ctx->cur], ctx->istage);
		CK(cudaMemcpyAsync(per_particle, ctx->istage, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	}
	rc = check_device_errors(ctx);
	if (rc) return rc;
	if (total) *total = ctx->slab ? (int64_t)llround(ctx->h_pinned[0]) : (int64_t)llround(ctx->h_pinned[0]) / 2;
	return SMD_OK;
}

// ------------------------------------------------------------------------------------------------ box moves
static int grow_cell_tables(smd_ctx *ctx, long long need)
{
	if (need > ctx->cellcap_limit) { ctx->err = "box grew beyond the cell-table limit of 64 Mi entries"; return SMD_ERR_UNSUPPORTED; }
	const long long cap = std::min<long long>(need + need / 4, ctx->cellcap_limit);
	CK(cudaStreamSynchronize(ctx->stream));
	int *cnt = nullptr, *st = nullptr, *cu = nullptr;
	unsigned long long *ss = nullptr;
	if (cudaMalloc(&cnt, (cap + 1) * sizeof(int)) != cudaSuccess || cudaMalloc(&st, (cap + 1) * sizeof(int)) != cudaSuccess ||
	    cudaMalloc(&cu, (cap + 1) * sizeof(int)) != cudaSuccess || cudaMalloc(&ss, (size_t)(2 * SCAN_BLOCKS) * sizeof(unsigned long long)) != cudaSuccess) {
		cudaFree(cnt); cudaFree(st); cudaFree(cu); cudaFree(ss);
		cudaGetLastError();
		ctx->err = "out of device memory growing the cell tables";
		return SMD_ERR_CUDA;
	}
	CK(cudaMemsetAsync(cnt, 0, (cap + 1) * sizeof(int), ctx->stream));   // the histogram is left zeroed by every build
	CK(cudaMemsetAsync(ss, 0, (size_t)(2 * SCAN_BLOCKS) * sizeof(unsigned long long), ctx->stream));
	cudaFree(ctx->count); cudaFree(ctx->start); cudaFree(ctx->cursor); cudaFree(ctx->scan_state);
	ctx->count = cnt; ctx->start = st; ctx->cursor = cu; ctx->scan_state = ss;
	ctx->cellcap = cap;
	ctx->cells_valid = false;
	ctx->hist_pending = false;     // (went with the old table)
	ctx->next_win_valid = false;
	return SMD_OK;
}

extern "C" int smd_rescale(smd_ctx *ctx, const double scale[3], const double new_box[3])
{
	if (!ctx) return SMD_ERR_ARG;
	ctx->du_ready = false;   // the particles change: block sums armed by smd_arm_dpotential no longer describe them
	REQUIRE(scale && new_box && ctx->particles_set, "bad call");
	CK(cudaSetDevice(ctx->device));
	Geom old = ctx->geom;
	const PairGeo old_pg = ctx->pgeo;
	set_geom(ctx, new_box);
	int rc = check_geom(ctx);
	long long total = (long long)ctx->geom.nc[0] * ctx->geom.nc[1] * ctx->geom.nc[2];
	// the grid outgrew the offset tables: CellOpt::resize reallocates (cellOpt.h:1525-1570), and so do we -- an accepted box
	// move is rare enough for a synchronisation
	if (!rc && total * std::max(2, ctx->geom.xs) > ctx->cellcap) rc = grow_cell_tables(ctx, total * std::max(2, ctx->geom.xs));
	if (!rc) rc = upload_acut(ctx);
	if (rc) {   // nothing has touched the particles yet: back to the old geometry on every error path
		ctx->geom = old; ctx->pgeo = old_pg;
		upload_acut(ctx);
		return rc;
	}
	LAUNCH(k_rescale, nblk(ctx->N, TPB), TPB, 0, cnt_of(ctx), ctx->pos[ctx->pcur], scale[0], scale[1], scale[2]);
	bool same_grid = old.nc[0] == ctx->geom.nc[0] && old.nc[1] == ctx->geom.nc[1] && old.nc[2] == ctx->geom.nc[2];
	retag_cells(ctx, !same_grid);
	return SMD_OK;
}

extern "C" int smd_mc_propose(const double box[3], double deltaLXY, double u_fluct, double new_box[3], double scale[3])
{
	if (!box || !new_box || !scale) return SMD_ERR_ARG;
	// MD.cpp:591-613
	double size[3] = {box[0], box[1], box[2]};
	double fl[3];
	fl[0] = deltaLXY * (2.0 * u_fluct - 1.0);
	fl[1] = fl[0];
	fl[2] = (size[0] * size[1]) / ((size[0] + fl[0]) * (size[1] + fl[1]));
	size[0] += fl[0]; size[1] += fl[1]; size[2] *= fl[2];
	for (int d = 0; d < 3; d++) { new_box[d] = size[d]; scale[d] = size[d] / box[d]; }
	return SMD_OK;
}

extern "C" int smd_mc_accept(double dU_terms_sum, double tension, const double box[3], const double new_box[3], double temperature,
                             double u_accept, int32_t *accepted, double *dU_total)
{
	if (!box || !new_box || !accepted) return SMD_ERR_ARG;
	double dPotential = dU_terms_sum;
	if (tension != 0) dPotential += tension * ((new_box[0] * new_box[1]) - (box[0] * box[1]));   // MD.cpp:677-678
	double D = exp(dPotential / temperature);                                                       // :683-687
	*accepted = (D >= u_accept || -dPotential <= 0) ? 1 : 0;                                        // :695
	if (dU_total) *dU_total = dPotential;
	return SMD_OK;
}

// the trial of MD.cpp:589-721 on the current configuration; `proposed`: box and scale already drawn (smd_step_mc)
static int mc_trial(smd_ctx *ctx, const double oldSize[3], const double size[3], const double aSize[3], double tension, double u_accept,
                    int32_t *accepted, double *dU_total, double box_out[3])
{
	double terms[SMD_NTERMS];
	int rc = smd_dpotential(ctx, aSize, terms);
	if (rc) return rc;
	// MD.cpp:615-675: pair first, then the molecules in file order (we sum by kind; FP64 sum order differs only)
	// MD.cpp:657-666 evaluates doNanoCoreDPotential / doBallDPotential and DROPS the result (no `dPotential+=`)
	double dPotential = 0;
	for (int t = 0; t < SMD_NTERMS; t++)
		if (t != SMD_TERM_BALL && t != SMD_TERM_NANOCORE) dPotential += terms[t];
	int32_t acc = 0;
	smd_mc_accept(dPotential, tension, oldSize, size, ctx->temperature, u_accept, &acc, &dPotential);
	if (acc) {
		rc = smd_rescale(ctx, aSize, size);
		if (rc) return rc;
	}
	if (accepted) *accepted = acc;
	if (dU_total) *dU_total = dPotential;
	if (box_out) for (int d = 0; d < 3; d++) box_out[d] = ctx->geom.box[d];
	return SMD_OK;
}

extern "C" int smd_mc_box_move(smd_ctx *ctx, double deltaLXY, double tension, double u_fluct, double u_accept, int32_t *accepted,
                               double *dU_total, double box_out[3])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(!ctx->slab, "slab: use smd_mc_propose / smd_dpotential / all-reduce / smd_mc_accept / smd_rescale");
	double oldSize[3] = {ctx->geom.box[0], ctx->geom.box[1], ctx->geom.box[2]};
	double size[3], aSize[3];
	smd_mc_propose(oldSize, deltaLXY, u_fluct, size, aSize);
	return mc_trial(ctx, oldSize, size, aSize, tension, u_accept, accepted, dU_total, box_out);
}

// The NEXT smd_step call lets the pair kernel of its last step also sum the pair dPotential of the box scaling `scale`
// (k_pair_force2 EMODE 3); the first smd_dpotential call for that same scale afterwards, with the particles untouched in
// between, takes the pair term from there instead of running a pass of its own.  Returns SMD_OK whether or not the fast
// path applies (asymmetric tables, external noise, ...: nothing is armed and smd_dpotential works as always).
extern "C" int smd_arm_dpotential(smd_ctx *ctx, const double scale[3])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(scale, "null scale");
	ctx->du_ready = false;
	ctx->du_for_last = false;
	const bool fuse = can_fuse(ctx) && ctx->tables_symmetric && !ctx->force_onephase_energy &&
	                  !ctx->no_du_fuse && ctx->desc.noise != SMD_NOISE_EXTERNAL;
	if (!fuse) return SMD_OK;
	CK(cudaSetDevice(ctx->device));
	const size_t need = (size_t)nblk(ctx->N, PAIR_SPLIT_NP);   // (either pair engine: the smaller block)
	if (ctx->du_partials_n < need) {   // block sums of its own: nothing else may overwrite them before they are used
		if (ctx->du_partials) cudaFree(ctx->du_partials);
		ctx->du_partials = nullptr; ctx->du_partials_n = 0;
		CK(cudaMalloc(&ctx->du_partials, need * sizeof(double)));
		ctx->du_partials_n = need;
	}
	EnergyArgs en;
	en.sx = scale[0]; en.sy = scale[1]; en.sz = scale[2];
	en.partials = ctx->du_partials; en.uC = ctx->uC; en.utab = ctx->utab;
	double grow = 0;   // as in energy_terms<2>: the most the scaling moves r^2 across a cutoff, relative
	for (double sc : {scale[0], scale[1], scale[2]}) grow = std::max(grow, std::max(fabs(sc * sc - 1.0), fabs(1.0 / (sc * sc) - 1.0)));
	en.extra32 = nextafterf((float)(1.01 * grow * ctx->geom.rc2 + 1e-7), INFINITY);
	ctx->du_en = en;
	ctx->du_for_last = true;
	return SMD_OK;
}

extern "C" int smd_step_mc(smd_ctx *ctx, int64_t first_step, int32_t nsteps, double deltaLXY, double tension, double u_fluct, double u_accept,
                           int32_t *accepted, double *dU_total, double box_out[3])
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(!ctx->slab, "slab: use smd_arm_dpotential / smd_step and smd_mc_propose / smd_dpotential / all-reduce / smd_mc_accept / smd_rescale");
	double oldSize[3] = {ctx->geom.box[0], ctx->geom.box[1], ctx->geom.box[2]};
	double size[3], aSize[3];
	smd_mc_propose(oldSize, deltaLXY, u_fluct, size, aSize);
	// The trial sees the positions of the last step's force evaluation (Verlet::second only moves velocities), so that
	// step's pair kernel can also sum the pair dPotential: one pass over the pairs instead of two.
	int rc = SMD_OK;
	if (nsteps > 0) {
		if ((rc = smd_arm_dpotential(ctx, aSize))) return rc;
		if ((rc = smd_step(ctx, first_step, nsteps))) return rc;
	}
	return mc_trial(ctx, oldSize, size, aSize, tension, u_accept, accepted, dU_total, box_out);
}

// ------------------------------------------------------------------------------------------------ read back
extern "C" int smd_get_particles(smd_ctx *ctx, double *xyz, int32_t *type, double *vel)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(ctx->particles_set, "smd_set_particles first");
	REQUIRE(!ctx->slab, "slab: use smd_slab_get_local");
	CK(cudaSetDevice(ctx->device));
	int N = ctx->N;
	double *sx = ctx->stage, *sv = ctx->stage + 3 * (size_t)ctx->cap;
	LAUNCH(k_export_particles, nblk(N, TPB), TPB, 0, N, ctx->cap, ctx->pos[ctx->pcur], ctx->vel[ctx->cur], ctx->gid[ctx->cur], xyz ? sx : nullptr,
	       type ? ctx->istage : nullptr, vel ? sv : nullptr);
	if (xyz) CK(cudaMemcpyAsync(xyz, sx, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	if (type) CK(cudaMemcpyAsync(type, ctx->istage, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	if (vel) CK(cudaMemcpyAsync(vel, sv, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	return check_device_errors(ctx);
}

extern "C" int smd_get_forces(smd_ctx *ctx, double *acc)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(acc && ctx->particles_set, "bad call");
	REQUIRE(!ctx->slab, "slab: use smd_slab_get_local");
	CK(cudaSetDevice(ctx->device));
	int N = ctx->N;
	LAUNCH(k_export_soa3, nblk(N, TPB), TPB, 0, N, ctx->cap, ctx->acc, ctx->gid[ctx->cur], ctx->stage);
	CK(cudaMemcpyAsync(acc, ctx->stage, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	return check_device_errors(ctx);
}

#ifdef SMD_PHASE_CLOCKS
extern "C" int smd_phase_clocks(smd_ctx *ctx, unsigned long long *out16, int reset)
{
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaMemcpyFromSymbol(out16, g_pc, 16 * sizeof(unsigned long long)));
	if (reset) { unsigned long long z[16] = {0}; CK(cudaMemcpyToSymbol(g_pc, z, sizeof z)); }
	return SMD_OK;
}
#endif

// smd_timeline: device-side time stamps of the kernels of one MD step (see TlScope)
extern "C" int smd_timeline(smd_ctx *ctx, int32_t enable)
{
	if (!ctx) return SMD_ERR_ARG;
	ctx->timeline = enable != 0;
	return SMD_OK;
}

extern "C" int smd_timeline_read(smd_ctx *ctx, double *us15)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(us15, "null argument");
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	unsigned long long t[48];
	CK(cudaMemcpyFromSymbol(t, g_tl, sizeof t));
	REQUIRE(t[0] != ~0ull && t[2] != 0ull, "smd_timeline_read: no instrumented step yet (smd_timeline(ctx, 1), then smd_step with nsteps >= 3 on a CHAIN-only system)");
	for (int k = 0; k < 15; k++) us15[k] = t[k] == ~0ull || t[k] == 0ull ? -1.0 : (double)(long long)(t[k] - t[0]) * 1e-3;
	return SMD_OK;
}

// ------------------------------------------------------------------------------------------------ observables
// a molecule the driver in charge parses and ignores (default case of MD.cpp:414-478: SOLID, OFFSET_BOUNDARY, RIGIDBEND,
// PULLBEAD; of MDsubstrate.cpp:213-262: SOLID, BALL, FLOATING_BASE, ZTORQUE, ZPOWERPOTENTIAL, NANOCORE); registered so that
// molecule k of smd_observe is molecule k of the file
extern "C" int smd_add_inert(smd_ctx *ctx, int32_t kind)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(kind == SMD_MOL_SOLID || kind == SMD_MOL_OFFSET_BOUNDARY || kind == SMD_MOL_RIGIDBEND || kind == SMD_MOL_PULLBEAD ||
	        kind == SMD_MOL_BALL || kind == SMD_MOL_FLOATING_BASE || kind == SMD_MOL_ZTORQUE || kind == SMD_MOL_ZPOWERPOTENTIAL ||
	        kind == SMD_MOL_NANOCORE, "smd_add_inert: only kinds one of the reference's drivers ignores");
	ctx->mol_order.push_back({kind, 0, 0});
	ctx->n_molecules++;
	return SMD_OK;
}

static const long long KE_BINS = 1ll << 21;      // 0.0001 per bin: kinetic energies up to 209 (|v| = 20 at unit mass)
static const int KE_OVERFLOW = 4096;
static const double KE_PARTITION = 0.0001;       // kEnergyDensityPartition, dataExtraction.h:390

static int obs_alloc(smd_ctx *ctx, int words)
{
	if (ctx->obs_words >= words) return SMD_OK;
	if (ctx->obs_buf) cudaFree(ctx->obs_buf);
	if (ctx->obs_host) cudaFreeHost(ctx->obs_host);
	ctx->obs_buf = nullptr; ctx->obs_host = nullptr; ctx->obs_words = 0;
	CK(cudaMalloc(&ctx->obs_buf, (size_t)words * 8));
	CK(cudaMemsetAsync(ctx->obs_buf, 0, (size_t)words * 8, ctx->stream));
	CK(cudaMallocHost(&ctx->obs_host, (size_t)words * 8));
	ctx->obs_words = words;
	return SMD_OK;
}

extern "C" int smd_msd_start(smd_ctx *ctx)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(ctx->particles_set && ctx->unw[0], "unwrapped positions are not tracked");
	REQUIRE(!ctx->slab, "slab: observables are not reduced per rank");
	CK(cudaSetDevice(ctx->device));
	const int N = ctx->N;
	if (!ctx->msd_start) CK(cudaMalloc(&ctx->msd_start, 3 * (size_t)N * sizeof(double)));
	LAUNCH(k_obs_msd_start, nblk(N, TPB), TPB, 0, N, ctx->cap, ctx->unw[ctx->cur], ctx->gid[ctx->cur], ctx->msd_start);
	return SMD_OK;
}

extern "C" int smd_observe(smd_ctx *ctx, uint32_t what, smd_observables *out, double *msd_sum, int64_t *msd_count, int32_t msd_cap)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(out && ctx->particles_set, "bad call");
	REQUIRE(!ctx->slab, "slab: observables are not reduced per rank");
	const bool msd = (what & SMD_OBS_MSD) != 0;
	REQUIRE(!msd || (ctx->msd_start && ctx->unw[0]), "SMD_OBS_MSD: smd_msd_start first (and track_unwrapped)");
	REQUIRE(!msd || (msd_sum && msd_count && msd_cap >= ctx->n_molecules), "SMD_OBS_MSD: one entry per molecule");
	CK(cudaSetDevice(ctx->device));
	const int N = ctx->N;
	const Particle *pos = ctx->pos[ctx->pcur];
	// slots of 8 bytes: [0..5] extent keys, [6] histogram overflow count, [8 ..] sums (one per launch below)
	int nsum = 0;
	nsum += (int)ctx->bonds.size() + 3 * (int)ctx->bends.size();
	nsum += (int)ctx->chains.size() + (int)ctx->bonds.size() + (int)ctx->bends.size() + (int)ctx->beads.size();
	int rc = obs_alloc(ctx, 8 + nsum + 8);
	if (rc) return rc;
	double *sums = reinterpret_cast<double *>(ctx->obs_buf + 8);
	int slot = 0;
	memset(out, 0, sizeof *out);

	std::vector<int> bond_slots, bend_slots;
	if (what & SMD_OBS_BONDS) {
		for (auto &b : ctx->bonds) {
			if (b.n <= 0) continue;
			const int nb = nblk(b.n, TPB);
			LAUNCH(k_obs_bond, nb, TPB, 0, b.n, b.d_ij, pos, ctx->slot_of, ctx->geom, ctx->partials);
			LAUNCH(k_final_sum, 1, 256, 0, nb, ctx->partials, sums, slot, 1.0);
			bond_slots.push_back(slot++);
			out->n_bond += b.n;
		}
		for (auto &b : ctx->bends) {
			if (b.n <= 0) continue;
			const int nb = nblk(b.n, TPB);
			REQUIRE(3 * (long long)nb <= MAX_PARTIALS, "BEND list too long for the partial-sum buffer");
			LAUNCH(k_obs_bend, nb, TPB, 0, b.n, b.d_ijk, pos, ctx->slot_of, ctx->geom, ctx->partials);
			for (int q = 0; q < 3; q++) LAUNCH(k_final_sum, 1, 256, 0, nb, ctx->partials + (size_t)q * nb, sums, slot + q, 1.0);
			bend_slots.push_back(slot);
			slot += 3;
			out->n_bend += b.n;
		}
	}
	if (what & (SMD_OBS_EXTENT | SMD_OBS_KE_HIST)) {
		ObsHist h = {nullptr, 0, nullptr, 0, nullptr};
		if (what & SMD_OBS_KE_HIST) {
			if (!ctx->ke_bins) {
				CK(cudaMalloc(&ctx->ke_bins, (size_t)KE_BINS * 8));
				CK(cudaMemsetAsync(ctx->ke_bins, 0, (size_t)KE_BINS * 8, ctx->stream));
				CK(cudaMalloc(&ctx->ke_overflow, (size_t)KE_OVERFLOW * 8));
			}
			CK(cudaMemsetAsync(ctx->obs_buf + 6, 0, 8, ctx->stream));
			h.bins = ctx->ke_bins; h.cap = KE_BINS; h.overflow = ctx->ke_overflow; h.overflow_cap = KE_OVERFLOW;
			h.n_overflow = reinterpret_cast<int *>(ctx->obs_buf + 6);
		}
		LAUNCH(k_obs_init, 1, 1, 0, ctx->obs_buf, ctx->geom.box[0], ctx->geom.box[1], ctx->geom.box[2]);
		LAUNCH(k_obs_particles, std::min(nblk(N, 256), 1184), 256, 0, cnt_of(ctx), ctx->cap, pos, ctx->vel[ctx->cur], ctx->gid[ctx->cur],
		       ctx->obs_buf, h, KE_PARTITION);
	}
	struct MsdSlot { int mol, slot; };
	std::vector<MsdSlot> msd_slots;
	if (msd) {
		const double *unw = ctx->unw[ctx->cur];
		for (int k = 0; k < ctx->n_molecules; k++) {
			msd_sum[k] = 0.0; msd_count[k] = 0;
			const smd_ctx::MolRef &r = ctx->mol_order[k];
			auto run = [&](int n, const int *idx, int first) -> int {
				if (n <= 0) return SMD_OK;
				const int nb = nblk(n, TPB);
				LAUNCH(k_obs_msd, nb, TPB, 0, n, idx, first, N, ctx->cap, unw, ctx->slot_of, ctx->msd_start, ctx->partials);
				LAUNCH(k_final_sum, 1, 256, 0, nb, ctx->partials, sums, slot, 1.0);
				msd_slots.push_back({k, slot++});
				msd_count[k] += n;
				return SMD_OK;
			};
			if (r.kind == SMD_MOL_CHAIN) {
				for (int j = r.first; j < r.first + r.count; j++)
					if ((rc = run(ctx->chains[j].nChains * ctx->chains[j].len, nullptr, ctx->chains[j].start))) return rc;
			} else if (r.kind == SMD_MOL_BOND) {
				if ((rc = run(2 * ctx->bonds[r.first].n, ctx->bonds[r.first].d_ij, 0))) return rc;
			} else if (r.kind == SMD_MOL_BEND) {
				if ((rc = run(3 * ctx->bends[r.first].n, ctx->bends[r.first].d_ijk, 0))) return rc;
			} else if (r.kind == SMD_MOL_BEAD) {
				if ((rc = run(ctx->beads[r.first].nOwn, ctx->beads[r.first].d_beads, 0))) return rc;
			}
		}
	}
	CK(cudaMemcpyAsync(ctx->obs_host, ctx->obs_buf, (size_t)(8 + slot) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	int novf = 0;
	if (what & SMD_OBS_KE_HIST) {
		// (the overflow list is read only when it is not empty: one more small copy, practically never)
	}
	rc = check_device_errors(ctx);   // synchronises
	if (rc) return rc;
	const double *hs = reinterpret_cast<const double *>(ctx->obs_host + 8);
	for (int sl : bond_slots) out->lbond_sum += hs[sl];
	for (int sl : bend_slots) { out->cos_bend_sum += hs[sl]; out->lbend_sum[0] += hs[sl + 1]; out->lbend_sum[1] += hs[sl + 2]; }
	if (what & (SMD_OBS_EXTENT | SMD_OBS_KE_HIST)) {
		auto unkey = [](unsigned long long k) {
			unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
			double v; memcpy(&v, &b, 8); return v;
		};
		for (int a = 0; a < 3; a++) { out->lo[a] = unkey(ctx->obs_host[a]); out->hi[a] = unkey(ctx->obs_host[3 + a]); }
	}
	if (what & SMD_OBS_KE_HIST) {
		novf = (int)(ctx->obs_host[6] & 0xffffffffull);
		REQUIRE(novf <= KE_OVERFLOW, "kinetic-energy histogram: more than 4096 particles beyond 0.5 v^2 = 209 (diverged?)");
		if (novf > 0) {
			std::vector<unsigned long long> ov((size_t)novf);
			CK(cudaMemcpy(ov.data(), ctx->ke_overflow, (size_t)novf * 8, cudaMemcpyDeviceToHost));
			for (unsigned long long b : ov) {
				bool found = false;
				for (auto &pr : ctx->ke_spill) if (pr.first == (long long)b) { pr.second++; found = true; break; }
				if (!found) ctx->ke_spill.push_back({(long long)b, 1});
			}
		}
	}
	for (auto &ms : msd_slots) msd_sum[ms.mol] += hs[ms.slot];
	out->n_molecules = msd ? ctx->n_molecules : 0;
	return SMD_OK;
}

extern "C" int smd_ke_histogram(smd_ctx *ctx, int64_t *counts, int64_t cap, int64_t *n_bins)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(n_bins, "bad call");
	*n_bins = 0;
	if (!ctx->ke_bins) return SMD_OK;   // nothing observed yet
	CK(cudaSetDevice(ctx->device));
	std::vector<unsigned long long> h((size_t)KE_BINS);
	CK(cudaMemcpyAsync(h.data(), ctx->ke_bins, (size_t)KE_BINS * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	long long top = -1;
	for (long long b = KE_BINS - 1; b >= 0; b--) if (h[(size_t)b]) { top = b; break; }
	for (auto &pr : ctx->ke_spill) top = std::max(top, pr.first);
	*n_bins = top + 1;
	if (counts) {
		const long long n = std::min<long long>(cap, top + 1);
		for (long long b = 0; b < n; b++) counts[b] = b < KE_BINS ? (int64_t)h[(size_t)b] : 0;
		for (auto &pr : ctx->ke_spill) if (pr.first < n) counts[pr.first] += pr.second;
	}
	return SMD_OK;
}

extern "C" int smd_get_unwrapped(smd_ctx *ctx, double *xyz)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(xyz && ctx->particles_set && ctx->unw[0], "unwrapped positions are not tracked");
	REQUIRE(!ctx->slab, "slab: unwrapped read-back is not exported");
	CK(cudaSetDevice(ctx->device));
	int N = ctx->N;
	LAUNCH(k_export_soa3, nblk(N, TPB), TPB, 0, N, ctx->cap, ctx->unw[ctx->cur], ctx->gid[ctx->cur], ctx->stage);
	CK(cudaMemcpyAsync(xyz, ctx->stage, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	return check_device_errors(ctx);
}

extern "C" int smd_host_alloc(void **ptr, size_t bytes)
{
	if (!ptr) return SMD_ERR_ARG;
	cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
	if (e != cudaSuccess) { *ptr = nullptr; g_create_error = std::string("smd_host_alloc: ") + cudaGetErrorString(e); return SMD_ERR_CUDA; }
	return SMD_OK;
}

extern "C" int smd_host_free(void *ptr)
{
	if (ptr) cudaFreeHost(ptr);
	return SMD_OK;
}

extern "C" int smd_snapshot(smd_ctx *ctx, double *xyz, double *vel, double *unwrapped, int64_t *ticket)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(ticket && ctx->particles_set, "bad call");
	REQUIRE(!ctx->slab, "slab: use smd_slab_get_local");
	REQUIRE(!unwrapped || ctx->unw[0], "unwrapped positions are not tracked");
	CK(cudaSetDevice(ctx->device));
	const int N = ctx->N;
	const size_t cap3 = 3 * (size_t)ctx->cap;
	if (!ctx->snap_stage) {
		CK(cudaMalloc(&ctx->snap_stage, 3 * cap3 * sizeof(double)));
		CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
		CK(cudaEventCreateWithFlags(&ctx->snap_gathered, cudaEventDisableTiming));
		for (int k = 0; k < 2; k++) CK(cudaEventCreateWithFlags(&ctx->snap_done[k], cudaEventDisableTiming | cudaEventBlockingSync));
	}
	// the gather buffer is free again once the copies of the previous ticket have read it
	const long long seq = ctx->snap_seq.load();
	if (seq > 0) CK(cudaStreamWaitEvent(ctx->stream, ctx->snap_done[(seq - 1) & 1], 0));
	double *sx = ctx->snap_stage, *sv = sx + cap3, *su = sv + cap3;
	if (xyz || vel)
		LAUNCH(k_export_particles, nblk(N, TPB), TPB, 0, N, ctx->cap, ctx->pos[ctx->pcur], ctx->vel[ctx->cur], ctx->gid[ctx->cur], xyz ? sx : nullptr,
		       (int *)nullptr, vel ? sv : nullptr);
	if (unwrapped) LAUNCH(k_export_soa3, nblk(N, TPB), TPB, 0, N, ctx->cap, ctx->unw[ctx->cur], ctx->gid[ctx->cur], su);
	CK(cudaEventRecord(ctx->snap_gathered, ctx->stream));
	CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->snap_gathered, 0));
	const size_t bytes = 3 * (size_t)N * sizeof(double);
	if (xyz) CK(cudaMemcpyAsync(xyz, sx, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
	if (vel) CK(cudaMemcpyAsync(vel, sv, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
	if (unwrapped) CK(cudaMemcpyAsync(unwrapped, su, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
	CK(cudaEventRecord(ctx->snap_done[seq & 1], ctx->copy_stream));
	*ticket = seq;
	ctx->snap_seq.store(seq + 1);
	return SMD_OK;
}

extern "C" int smd_snapshot_wait(smd_ctx *ctx, int64_t ticket)
{
	if (!ctx) return SMD_ERR_ARG;
	// called from the writer thread while the owner keeps enqueueing work: touches nothing but the ticket's event
	const long long seq = ctx->snap_seq.load();
	if (ticket < 0 || ticket >= seq || ticket + 2 < seq) return SMD_ERR_ARG;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return SMD_ERR_CUDA;
	return cudaEventSynchronize(ctx->snap_done[ticket & 1]) == cudaSuccess ? SMD_OK : SMD_ERR_CUDA;
}

extern "C" int smd_get_box(smd_ctx *ctx, double box[3])
{
	if (!ctx || !box) return SMD_ERR_ARG;
	for (int d = 0; d < 3; d++) box[d] = ctx->geom.box[d];
	return SMD_OK;
}

extern "C" int smd_get_cell_ids(smd_ctx *ctx, int32_t n_cells_xyz[3], int32_t *cell_key, int32_t *cell_rank)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(cell_key && cell_rank, "null output");
	REQUIRE(!ctx->slab, "slab: cell ids are not exported");
	int rc = ready(ctx);
	if (rc) return rc;
	if (!ctx->cells_valid) build_cells(ctx);
	int N = ctx->N;
	int *key = ctx->istage, *rank = ctx->istage + ctx->cap;
	LAUNCH(k_export_cells, nblk(N, TPB), TPB, 0, N, ctx->pos[ctx->pcur], ctx->gid[ctx->cur], ctx->start, cur_win(ctx), ctx->geom, key, rank);
	CK(cudaMemcpyAsync(cell_key, key, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaMemcpyAsync(cell_rank, rank, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	if (n_cells_xyz) for (int d = 0; d < 3; d++) n_cells_xyz[d] = ctx->geom.nc[d];
	return check_device_errors(ctx);
}

extern "C" int smd_device_ptr(smd_ctx *ctx, int32_t which, void **ptr, size_t *bytes)
{
	if (!ctx || !ptr) return SMD_ERR_ARG;
	size_t b = 0;
	switch (which) {
	case 0: *ptr = ctx->pos[ctx->pcur]; b = (size_t)ctx->N * sizeof(Particle); break;
	case 1: *ptr = ctx->vel[ctx->cur]; b = 3 * (size_t)ctx->cap * sizeof(double); break;
	case 2: *ptr = ctx->acc; b = 3 * (size_t)ctx->cap * sizeof(double); break;
	case 3: *ptr = ctx->gid[ctx->cur]; b = (size_t)ctx->N * sizeof(int); break;
	default: ctx->err = "unknown buffer"; return SMD_ERR_ARG;
	}
	if (bytes) *bytes = b;
	return SMD_OK;
}

extern "C" int smd_stream(smd_ctx *ctx, void **stream)
{
	if (!ctx || !stream) return SMD_ERR_ARG;
	*stream = (void *)ctx->stream;
	return SMD_OK;
}

extern "C" int smd_fp64_peak(smd_ctx *ctx, double *fma_tflops, double *muladd_tflops)
{
	if (!ctx) return SMD_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, ctx->device));
	int blocks = prop.multiProcessorCount * 8, iters = 4096;
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	double res[2] = {0, 0};
	for (int mode = 0; mode < 2; mode++) {
		float best = 1e30f;
		for (int rep = 0; rep < 4; rep++) {
			CK(cudaEventRecord(e0, ctx->stream));
			if (mode == 0) k_fp64_peak<1><<<blocks, 256, 0, ctx->stream>>>(iters, 0.999999, 1e-7, ctx->scalars);
			else k_fp64_peak<0><<<blocks, 256, 0, ctx->stream>>>(iters, 0.999999, 1e-7, ctx->scalars);
			ctx->launches++;
			CK(cudaEventRecord(e1, ctx->stream));
			CK(cudaEventSynchronize(e1));
			float ms = 0;
			CK(cudaEventElapsedTime(&ms, e0, e1));
			if (rep > 0 && ms < best) best = ms;
		}
		double flop = (double)blocks * 256.0 * iters * 8.0 * 2.0;   // mul + add per element, fused or not
		res[mode] = flop / (best * 1e-3) * 1e-12;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	if (fma_tflops) *fma_tflops = res[0];
	if (muladd_tflops) *muladd_tflops = res[1];
	return SMD_OK;
}

extern "C" int smd_stats(smd_ctx *ctx, int64_t *kernel_launches, int64_t *rebuilds)
{
	if (!ctx) return SMD_ERR_ARG;
	if (kernel_launches) *kernel_launches = ctx->launches;
	if (rebuilds) *rebuilds = ctx->rebuilds;
	return SMD_OK;
}

// ------------------------------------------------------------------------------------------------ slab decomposition
extern "C" int smd_slab_columns(int32_t n_cols, int32_t nranks, int32_t rank, int32_t *col_lo, int32_t *col_hi)
{
	if (nranks <= 0 || rank < 0 || rank >= nranks || n_cols <= 0 || !col_lo || !col_hi) return SMD_ERR_ARG;
	*col_lo = (int32_t)(((long long)rank * n_cols) / nranks);
	*col_hi = (int32_t)(((long long)(rank + 1) * n_cols) / nranks);
	return SMD_OK;
}

extern "C" int smd_slab_select(const double box[3], double cutoff, int32_t nranks, int32_t rank, int32_t n, const double *xyz,
                               int32_t *flags)
{
	if (!box || !xyz || !flags || cutoff <= 0) return SMD_ERR_ARG;
	int nc = (int)(box[0] / cutoff);              // cellOpt.h:194
	double cs = box[0] / nc;                      // cellOpt.h:207
	int32_t lo, hi;
	int rc = smd_slab_columns(nc, nranks, rank, &lo, &hi);
	if (rc) return rc;
	const int W = hi - lo, H = SMD_SLAB_HALO;
	for (int i = 0; i < n; i++) {
		int cx = (int)(xyz[3 * (size_t)i] / cs);  // cellOpt.h:532
		if (cx >= nc) cx -= 1;                    // cellOpt.h:537
		int rel = cx - lo;
		if (rel < 0) rel += nc;
		flags[i] = rel < W ? 1 : ((nranks > 1 && (rel < W + H || rel >= nc - H)) ? 2 : 0);
	}
	return SMD_OK;
}

extern "C" int smd_slab_recv_buffer(smd_ctx *ctx, int32_t side, void **ptr, size_t *bytes)
{
	if (!ctx || !ptr) return SMD_ERR_ARG;
	REQUIRE(ctx->slab && (side == 0 || side == 1), "not a slab context / bad side");
	*ptr = ctx->recv_base[side];
	if (bytes) *bytes = ctx->recv_bytes;
	return SMD_OK;
}

extern "C" int smd_slab_ipc_handle(smd_ctx *ctx, int32_t side, void *handle64)
{
	if (!ctx || !handle64) return SMD_ERR_ARG;
	REQUIRE(ctx->slab && (side == 0 || side == 1), "not a slab context / bad side");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
	CK(cudaSetDevice(ctx->device));
	cudaIpcMemHandle_t h;
	CK(cudaIpcGetMemHandle(&h, ctx->recv_base[side]));
	memcpy(handle64, &h, sizeof h);
	return SMD_OK;
}

extern "C" int smd_slab_connect_ptr(smd_ctx *ctx, int32_t dir, void *peer_buffer)
{
	if (!ctx || !peer_buffer) return SMD_ERR_ARG;
	REQUIRE(ctx->slab && (dir == 0 || dir == 1), "not a slab context / bad direction");
	ctx->comm.send[dir] = (char *)peer_buffer;
	ctx->peer_set[dir] = true;
	return SMD_OK;
}

extern "C" int smd_slab_connect_ipc(smd_ctx *ctx, int32_t dir, const void *handle64)
{
	if (!ctx || !handle64) return SMD_ERR_ARG;
	REQUIRE(ctx->slab && (dir == 0 || dir == 1), "not a slab context / bad direction");
	CK(cudaSetDevice(ctx->device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, sizeof h);
	void *p = nullptr;
	CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	ctx->ipc_opened[dir] = p;
	return smd_slab_connect_ptr(ctx, dir, p);
}

extern "C" int smd_slab_exchange_send(smd_ctx *ctx)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(ctx->slab, "not a slab context");
	REQUIRE(ctx->peer_set[0] && ctx->peer_set[1], "slab: connect both neighbours first (smd_slab_connect_*)");
	REQUIRE(!ctx->exch_pending, "slab: previous exchange not received yet");
	CK(cudaSetDevice(ctx->device));
	ProfScope ps(ctx, SMD_PHASE_EXCHANGE);
	ctx->xseq++;
	LAUNCH(k_slab_pack, nblk(ctx->N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->vel[ctx->cur], ctx->unw[ctx->cur],
	       ctx->gid[ctx->cur], ctx->geom, ctx->comm, ctx->xseq, ctx->errflag, ctx->slot_of);
	ctx->exch_pending = true;
	ctx->cells_valid = false;
	return SMD_OK;
}

extern "C" int smd_slab_exchange_recv(smd_ctx *ctx)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(ctx->slab && ctx->exch_pending, "slab: no exchange in flight");
	CK(cudaSetDevice(ctx->device));
	ProfScope ps(ctx, SMD_PHASE_EXCHANGE);
	// A few blocks only (grid-stride over <= 2 * capmsg entries): the kernel SPINS on the neighbours' headers, and when
	// several ranks share one device (the single-process test harness) their waiting kernels must never fill the machine
	// and starve the pack kernel they wait for (four ranks x 296 blocks x 256 threads did exactly that, intermittently).
	// One slab context on this device (the production layout: one process per GPU): nothing else can be starved, one block
	// per SM takes the entries in one trip.
	const bool alone = g_slab_ctx_on_device[ctx->device & 63] <= 1;
	int blocks = std::min(nblk(2ll * ctx->comm.capmsg, 256), alone ? 148 : 24);
	long long spin_limit = 20000000000ll;   // ~10 s of SM clocks: a neighbour that never sends is reported, not waited for
	LAUNCHP(k_slab_unpack, blocks, 256, 0, cnt_of(ctx), ctx->dN + 1, ctx->cap, ctx->pos[ctx->pcur], ctx->vel[ctx->cur], ctx->unw[ctx->cur],
	       ctx->gid[ctx->cur], ctx->comm, ctx->xseq, ctx->errflag, spin_limit, ctx->geom,
	       ctx->hist_pending ? BinArgs{next_win(ctx), ctx->count, ctx->cellOfSlot} : BinArgs{nullptr, nullptr, nullptr});
	ctx->exch_pending = false;
	ctx->ext_valid = true;
	return SMD_OK;
}

extern "C" int smd_slab_set_local(smd_ctx *ctx, int32_t n, const int32_t *gid, const double *xyz, const int32_t *type, const double *vel)
{
	if (!ctx) return SMD_ERR_ARG;
	ctx->du_ready = false;   // the particles change: block sums armed by smd_arm_dpotential no longer describe them
	REQUIRE(ctx->slab, "not a slab context");
	REQUIRE(n >= 0 && (n == 0 || (gid && xyz && type)), "null arrays");
	REQUIRE(n <= ctx->cap, "slab: local particle capacity exceeded at load (desc.reserved[0])");
	REQUIRE(ctx->peer_set[0] && ctx->peer_set[1], "slab: connect both neighbours first (smd_slab_connect_*)");
	REQUIRE(!ctx->exch_pending, "slab: previous exchange not received yet");
	CK(cudaSetDevice(ctx->device));
	// (global index, type and position of every particle are checked by the import kernel -- a host loop over 10^6 particles
	// cost milliseconds on the end-to-end path of every per-rank upload)
	double *sx = ctx->stage, *sv = ctx->stage + 3 * (size_t)ctx->cap;
	int *dg = ctx->istage + ctx->cap;
	if (!ctx->import_bad) CK(cudaMalloc(&ctx->import_bad, 2 * sizeof(int)));
	CK(cudaMemsetAsync(ctx->import_bad, 0x7f, 2 * sizeof(int), ctx->stream));
	CK(cudaMemcpyAsync(sx, xyz, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(ctx->istage, type, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(dg, gid, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	if (vel) CK(cudaMemcpyAsync(sv, vel, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemsetAsync(ctx->slot_of, 0xff, (size_t)ctx->n_global * sizeof(int), ctx->stream));
	int nn[2] = {n, n};
	CK(cudaMemcpyAsync(ctx->dN, nn, sizeof nn, cudaMemcpyHostToDevice, ctx->stream));
	ctx->cur = 0;
	ctx->pcur = 0;
	ctx->ext_valid = false;
	if (n > 0)
		LAUNCH(k_import_particles, nblk(n, TPB), TPB, 0, n, ctx->cap, sx, ctx->istage, vel ? sv : nullptr, ctx->pos[0], ctx->vel[0], ctx->unw[0],
		       ctx->gid[0], ctx->slot_of, dg, ctx->geom, ctx->nT, ctx->import_bad, ctx->n_global);
	int *h_bad = reinterpret_cast<int *>(ctx->h_pinned + 60);
	CK(cudaMemcpyAsync(h_bad, ctx->import_bad, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaMemsetAsync(ctx->acc, 0, 3 * (size_t)ctx->cap * sizeof(double), ctx->stream));
	ctx->acc_live = false;
	retag_cells(ctx);
	CK(cudaStreamSynchronize(ctx->stream));   // host buffers may be reused by the caller
	if (h_bad[0] != 0x7f7f7f7f || h_bad[1] != 0x7f7f7f7f) {
		ctx->particles_set = false;
		CK(cudaMemsetAsync(ctx->errflag, 0, sizeof(int), ctx->stream));
		if (h_bad[1] != 0x7f7f7f7f) { ctx->err = "global particle index out of range"; return SMD_ERR_ARG; }
		if (h_bad[0] % 4 == 3) { ctx->err = "particle type out of range"; return SMD_ERR_ARG; }
		ctx->err = "position of a particle is out of bounds.";   // system.h:452-469
		return SMD_ERR_CELL;
	}
	ctx->particles_set = true;
	// ghosts come from the neighbours (and strays go to them) through the ordinary exchange; received by the next
	// force evaluation
	return smd_slab_exchange_send(ctx);
}

extern "C" int smd_slab_capacity(smd_ctx *ctx, int32_t *capacity)
{
	if (!ctx || !capacity) return SMD_ERR_ARG;
	REQUIRE(ctx->slab, "not a slab context");
	*capacity = ctx->cap;
	return SMD_OK;
}

extern "C" int smd_slab_get_local(smd_ctx *ctx, int32_t *n, int32_t *gid, double *xyz, int32_t *type, double *vel, double *acc)
{
	if (!ctx || !n) return SMD_ERR_ARG;
	REQUIRE(ctx->slab && ctx->particles_set, "not a loaded slab context");
	REQUIRE(!ctx->exch_pending, "slab: read-back between smd_step_begin and smd_step_end");
	CK(cudaSetDevice(ctx->device));
	size_t cap = ctx->cap;
	if (!ctx->export_i) {   // export buffers: allocated once (cudaMalloc / cudaFree per read-back cost milliseconds and a device-wide wait)
		CK(cudaMalloc(&ctx->export_i, 2 * cap * sizeof(int)));
		CK(cudaMalloc(&ctx->export_d, 9 * cap * sizeof(double)));
	}
	int *d_i = ctx->export_i;
	double *d_d = ctx->export_d;
	CK(cudaMemsetAsync(ctx->d_export_counter, 0, sizeof(int), ctx->stream));
	LAUNCH(k_slab_export, nblk(ctx->N, TPB), TPB, 0, cnt_of(ctx), ctx->cap, ctx->pos[ctx->pcur], ctx->vel[ctx->cur], ctx->acc, ctx->gid[ctx->cur],
	       ctx->d_export_counter, d_i, d_d, d_d + 3 * cap, d_d + 6 * cap, d_i + cap);
	int *h_cnt = reinterpret_cast<int *>(ctx->h_pinned + 62);
	CK(cudaMemcpyAsync(h_cnt, ctx->d_export_counter, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	const int cnt = *h_cnt;
	// (asynchronous on the stream: with page-locked destinations the five copies overlap their set-up, one wait at the end)
	if (gid) CK(cudaMemcpyAsync(gid, d_i, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	if (type) CK(cudaMemcpyAsync(type, d_i + cap, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	if (xyz) CK(cudaMemcpyAsync(xyz, d_d, 3 * (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	if (vel) CK(cudaMemcpyAsync(vel, d_d + 3 * cap, 3 * (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	if (acc) CK(cudaMemcpyAsync(acc, d_d + 6 * cap, 3 * (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	*n = cnt;
	return check_device_errors(ctx);
}

extern "C" int smd_slab_counts(smd_ctx *ctx, int32_t *n_local, int32_t *n_owned)
{
	if (!ctx) return SMD_ERR_ARG;
	REQUIRE(ctx->slab && ctx->particles_set, "not a loaded slab context");
	CK(cudaSetDevice(ctx->device));
	int nn[2] = {0, 0};
	CK(cudaMemcpyAsync(nn, ctx->dN, sizeof nn, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (n_local) *n_local = nn[0];
	if (n_owned) {
		int32_t k = 0;
		int rc = smd_slab_get_local(ctx, &k, nullptr, nullptr, nullptr, nullptr, nullptr);
		if (rc) return rc;
		*n_owned = k;
	}
	return check_device_errors(ctx);
}

#ifdef SMD_EXP_TIMING
// experiment builds only (not part of the ABI)
extern "C" int smd_exp_pair_timing(unsigned long long out[8], int reset)
{
	cudaDeviceSynchronize();
	cudaMemcpyFromSymbol(out, g_pair_timing, 8 * sizeof(unsigned long long));
	if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(g_pair_timing, z, sizeof z); }
	return 0;
}
#endif
