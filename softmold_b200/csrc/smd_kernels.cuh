// Device kernels of the SoftMold MD timestep for sm_100a.  Included by smd_core.cu only.
// All FP64; compiled with --fmad=false (see smd_internal.cuh).  Reference citations are file:line relative to the
// reference root.
#pragma once
#include "smd_internal.cuh"

namespace smd {

constexpr int TPB = 128;          // threads per block for per-particle kernels
constexpr int SCAN_BLOCKS = 296;  // persistent grid of the single-pass cell-offset scan (two blocks per SM)
constexpr int SCAN_TPB = 256;

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// deterministic block sum (result valid in thread 0); blockDim.x multiple of 32, <= 1024
__device__ __forceinline__ double block_sum(double v)
{
	__shared__ double sh[32];
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	v = warp_sum(v);
	__syncthreads();
	if (lane == 0) sh[w] = v;
	__syncthreads();
	int nw = (blockDim.x + 31) >> 5;
	v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
	if (w == 0) v = warp_sum(v);
	return v;
}

// Programmatic dependent launch (the kernels of the step loop are launched with the stream-serialization attribute when
// nothing is recorded between them): wait for the grid before this one to complete and flush -- a no-op under a plain
// launch --, then let the grid after this one become resident on whatever this one leaves free.  Every kernel launched
// with the attribute calls this first thing (k_chain_kick orders itself by completion words instead), so completion is
// transitive along the chain.
__device__ __forceinline__ void pdl_prologue()
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Device-side time stamps (%globaltimer) of the step's kernels -- when the first block of a kernel starts working (after its
// griddepcontrol.wait), when its last block starts and when its last block ends -- for smd_timeline / bench.py's
// `device_timeline_us` / tools/timeline.py.  Under programmatic dependent launches the kernels of a step overlap, which no
// event pair can show (and event pairs switch the overlap off).  The switch lives in constant memory: off, it costs one
// uniform constant load and a predicate per block.
__device__ unsigned long long g_tl[48];
__constant__ int c_tl_on;
__device__ __forceinline__ unsigned long long tl_now()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
struct TlScope {
	int id;
	bool on;
	__device__ __forceinline__ TlScope(int i) : id(i), on(threadIdx.x == 0 && c_tl_on != 0)
	{
		if (on) { const unsigned long long t = tl_now(); atomicMin(&g_tl[3 * id], t); atomicMax(&g_tl[3 * id + 1], t); }
	}
	__device__ __forceinline__ ~TlScope() { if (on) atomicMax(&g_tl[3 * id + 2], tl_now()); }
};
__global__ void k_tl_reset()
{
	for (int k = 0; k < 16; k++) { g_tl[3 * k] = ~0ull; g_tl[3 * k + 1] = 0ull; g_tl[3 * k + 2] = 0ull; }
}
#define SMD_TL(id) TlScope tl_scope_(id)
#ifdef SMD_PHASE_CLOCKS
// debug builds only (tools/phase_clocks.py): cycles per warp between the phase marks of k_pair_force2, summed over warps
__device__ unsigned long long g_pc[16];
#define SMD_PC(k) do { if ((threadIdx.x & 31) == 0) { const long long t_ = clock64(); atomicAdd(&g_pc[k], (unsigned long long)(t_ - pc_t)); pc_t = t_; } } while (0)
#define SMD_PC_INIT long long pc_t = clock64(); if (threadIdx.x == 0) atomicAdd(&g_pc[15], 1ull)
#else
#define SMD_PC(k)
#define SMD_PC_INIT
#endif

__device__ __forceinline__ Particle load_particle(const Particle *p)
{
	// two 16-byte loads of one aligned 32-byte record (one sector)
	const double2 *q = reinterpret_cast<const double2 *>(p);
	double2 a = q[0], b = q[1];
	Particle r;
	r.x = a.x; r.y = a.y; r.z = b.x;
	long long w = __double_as_longlong(b.y);
	r.type = (int)(w & 0xffffffffll);
	r.cell = (unsigned)((unsigned long long)w >> 32);
	return r;
}

__device__ __forceinline__ void store_particle(Particle *p, const Particle &r)
{
	double2 *q = reinterpret_cast<double2 *>(p);
	long long w = ((long long)(unsigned long long)r.cell << 32) | (unsigned)r.type;
	q[0] = make_double2(r.x, r.y);
	q[1] = make_double2(r.z, __longlong_as_double(w));
}

// ------------------------------------------------------------------------------------------------ Philox4x32-10
// Counter-based noise keyed on (seed, step, particle): spec in DESIGN.md, restated in oracle/oracle.c.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4])
{
#pragma unroll
	for (int r = 0; r < 10; r++) {
		uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
		uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
		uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
		c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
	return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void philox_uniform3(uint64_t seed, uint64_t step, uint32_t id, double u[3])
{
	uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ 0x5D0F7A3Bu;
	uint32_t w[4], x[4];
	philox4x32_10(id, (uint32_t)step, (uint32_t)(step >> 32), 0u, k0, k1, w);
	philox4x32_10(id, (uint32_t)step, (uint32_t)(step >> 32), 1u, k0, k1, x);
	u[0] = u53(w[0], w[1]);
	u[1] = u53(w[2], w[3]);
	u[2] = u53(x[0], x[1]);
}

// ------------------------------------------------------------------------------------------------ integrator
// Verlet::first, algorithms/verlet.h:288-356: v += a*dt/2 ; p += v*dt ; aP += v*dt (type != 0 only) ; wrap with
// strict > L and < 0 for every particle.  In place, slot order.
// It also does the first half of CellOpt::build for the new positions (cellOpt.h:530-556): the cell coordinates of
// the reference grid are packed into the record and the bounding box of occupied cells is accumulated for the
// counting sort that follows (tag_cell below).
__device__ __forceinline__ void tag_cell(Particle &p, const Geom &g, int *bbox, int *errflag, bool live);
__device__ __forceinline__ int bin_particle(const Particle &p, bool live, const Geom &g, const BinArgs &b, int *errflag);

__global__ void __launch_bounds__(TPB) k_verlet_first(Cnt cnt, int cap, Particle *pos, double *vel, const double *acc, double *unw,
                                                      Geom g, double dt, int *bbox, int *errflag, const int *gid, BinArgs bin)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	bool live = s < cnt.get();
	if (g.slab && live) live = !(gid[s] & GID_GHOST);   // ghosts are replaced by the exchange that follows
	Particle p;
	if (live) {
		p = load_particle(pos + s);
		double h = 0.5 * dt;
		if (p.type != 0) {
			double vx = vel[s], vy = vel[cap + s], vz = vel[2 * cap + s];
			vx += (acc[s] * h); vy += (acc[cap + s] * h); vz += (acc[2 * cap + s] * h);
			vel[s] = vx; vel[cap + s] = vy; vel[2 * cap + s] = vz;
			p.x += vx * dt; p.y += vy * dt; p.z += vz * dt;
			if (unw) {
				unw[s] += vx * dt; unw[cap + s] += vy * dt; unw[2 * cap + s] += vz * dt;
			}
		}
		if (p.x > g.box[0]) p.x -= g.box[0];
		if (p.x < 0) p.x += g.box[0];
		if (p.y > g.box[1]) p.y -= g.box[1];
		if (p.y < 0) p.y += g.box[1];
		if (p.z > g.box[2]) p.z -= g.box[2];
		if (p.z < 0) p.z += g.box[2];
	}
	tag_cell(p, g, bbox, errflag, live);
	if (live) store_particle(pos + s, p);
	if (bin.count) {
		const int local = bin_particle(p, live, g, bin, errflag);
		if (s < cnt.get()) bin.cellOfSlot[s] = local;
	}
}

// Verlet::second, algorithms/verlet.h:463-477
__global__ void __launch_bounds__(TPB) k_verlet_second(Cnt cnt, int cap, const Particle *pos, double *vel, const double *acc, double dt)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= cnt.get()) return;
	if (pos[s].type == 0) return;
	double h = 0.5 * dt;
	vel[s] += (acc[s] * h);
	vel[cap + s] += (acc[cap + s] * h);
	vel[2 * cap + s] += (acc[2 * cap + s] * h);
}

__global__ void __launch_bounds__(TPB) k_zero3(Cnt cnt, int cap, double *a)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= cnt.get()) return;
	a[s] = 0; a[cap + s] = 0; a[2 * cap + s] = 0;
}

// Langevin::compute scalar-gamma branch, algorithms/langevin.h:284-331: a += -g v + sigma (2u-1)
__global__ void __launch_bounds__(TPB) k_langevin(Cnt cnt, int cap, const double *vel, double *acc, const int *gid, double gamma,
                                                  double sigma, uint64_t seed, uint64_t step, const double *ext_noise)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= cnt.get()) return;
	int id = gid[s];
	if (id & GID_GHOST) return;
	double u[3];
	if (ext_noise) {
		u[0] = ext_noise[3 * id]; u[1] = ext_noise[3 * id + 1]; u[2] = ext_noise[3 * id + 2];
	} else {
		philox_uniform3(seed, step, (uint32_t)id, u);
	}
#pragma unroll
	for (int c = 0; c < 3; c++) {
		double psi = 2.0 * u[c] - 1.0;
		acc[c * cap + s] += (-gamma * vel[c * cap + s] + sigma * psi);
	}
}

// Kinetic::compute, algorithms/dataCollection.h:666-674: sum (vx^2+vy^2+vz^2)/2
__global__ void __launch_bounds__(256) k_kinetic(Cnt cnt, int cap, const double *vel, const int *gid, double *partials)
{
	double e = 0;
	const int N = cnt.get();
	for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < N; s += gridDim.x * blockDim.x) {
		if (gid[s] & GID_GHOST) continue;
		double vx = vel[s], vy = vel[cap + s], vz = vel[2 * cap + s];
		e += ((vx * vx + vy * vy + vz * vz) / 2.0);
	}
	e = block_sum(e);
	if (threadIdx.x == 0) partials[blockIdx.x] = e;
}

// final deterministic reduction of `n` partials into out[slot]
__global__ void __launch_bounds__(256) k_final_sum(int n, const double *partials, double *out, int slot, double factor)
{
	double e = 0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) e += partials[i];
	e = block_sum(e);
	if (threadIdx.x == 0) out[slot] = e * factor;
}

// smd_dpotential_device: the per-launch sums of an energy evaluation folded into one value per term, on the device
struct SlotTerms { int n; signed char term[64]; };
__global__ void k_fold_terms(SlotTerms st, const double *scalars, double *terms, int nterms)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	for (int t = 0; t < nterms; t++) terms[t] = 0.0;
	for (int k = 0; k < st.n; k++) terms[st.term[k]] += scalars[k];   // the host's order (energy_terms): bit-identical sums
}

// accepted box move: p *= aSize (MD.cpp:697-707)
__global__ void __launch_bounds__(TPB) k_rescale(Cnt cnt, Particle *pos, double sx, double sy, double sz)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= cnt.get()) return;
	Particle p = load_particle(pos + s);
	p.x *= sx; p.y *= sy; p.z *= sz;
	store_particle(pos + s, p);
}

// ------------------------------------------------------------------------------------------------ cell list
// CellOpt::build, cellOpt.h:510-710.  The reference grid is nc = int(L/rc), cell size L/nc, key from int(p/cs) with
// the upper-edge clamp (:530-556).  We keep a dense offset table only over the occupied window of that grid.

__device__ __forceinline__ bool cell_coords(const Particle &p, const Geom &g, int &cx, int &cy, int &cz)
{
	cx = (int)(p.x / g.cs[0]); cy = (int)(p.y / g.cs[1]); cz = (int)(p.z / g.cs[2]);
	cx = (cx >= g.nc[0]) ? cx - 1 : cx;   // "correction for older systems", cellOpt.h:537-539
	cy = (cy >= g.nc[1]) ? cy - 1 : cy;
	cz = (cz >= g.nc[2]) ? cz - 1 : cz;
	// the reference would index out of bounds here (ERRORS_ENABLED: throw 0, cellOpt.h:541-552)
	return !(cx < 0 || cy < 0 || cz < 0 || cx >= g.nc[0] || cy >= g.nc[1] || cz >= g.nc[2]) &&
	       p.x == p.x && p.y == p.y && p.z == p.z;
}

// cell coordinates of p in the reference grid -> packed into the record; bounding box of occupied cells -> bbox[6]
// (warp-reduced, then one atomicMin/Max per warp).  Must be called by all 32 lanes.
__device__ __forceinline__ void tag_cell(Particle &p, const Geom &g, int *bbox, int *errflag, bool live)
{
	int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
	if (live) {
		int c[3];
		if (!cell_coords(p, g, c[0], c[1], c[2])) {
			atomicOr(errflag, ERR_OUT_OF_BOX);
			for (int d = 0; d < 3; d++) c[d] = min(max(c[d], 0), g.nc[d] - 1);
		}
		for (int d = 0; d < 3; d++) lo[d] = hi[d] = c[d];
		p.cell = pack_cell(c[0], c[1], c[2]);
	}
	if (g.slab) return;   // the window of a slab is fixed by its column range
	for (int d = 0; d < 3; d++) {
		int l = __reduce_min_sync(0xffffffffu, lo[d]), h = __reduce_max_sync(0xffffffffu, hi[d]);
		// the box of occupied cells hardly moves between steps: only touch the accumulators when they would change.
		// Plain (L1-cached) loads on purpose: the extremes only ever grow during a pass, so a stale copy can at worst
		// cause a redundant atomic, while volatile loads of one address by every warp of the grid serialise in a
		// single L2 slice (that was 15 us of the 20 us of the integrator kernel).
		if ((threadIdx.x & 31) == 0 && l != INT_MAX) {
			if (l < bbox[d]) atomicMin(bbox + d, l);
			if (h > bbox[3 + d]) atomicMax(bbox + 3 + d, h);
		}
	}
}

// stand-alone tagging pass (after set_particles / a box move; the steady state does it inside k_verlet_first)
__global__ void __launch_bounds__(TPB) k_tag_cells(Cnt cnt, Particle *pos, Geom g, int *bbox, int *errflag, const int *gid)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	bool live = s < cnt.get();   // slab mode: ghosts are re-tagged too (an accepted box move scales them like their owners)
	(void)gid;
	Particle p;
	if (live) p = load_particle(pos + s);
	tag_cell(p, g, bbox, errflag, live);
	if (live) store_particle(pos + s, p);
}

// The offset table covers only the occupied window of the reference grid: the exact bounding box of the cells
// tagged above, plus one empty cell on each side so that stencil look-ups of boundary cells stay inside the table.
// Every kernel of the build derives it from bbox[] with this one function; k_scan3 publishes it in win[] for the
// kernels that run after the build and re-arms bbox[].
struct Window { int org[3], dim[3]; int ncells; int fd0; };   // fd0 = dim[0] * xs: x extent of the offset table in slices

// window-local x index of cell column cx.  A slab's window starts `halo` columns left of its first owned column and
// may run across the periodic boundary, so the index wraps modulo nc[0]; for a single-GPU window (a sub-box of the
// grid) a column outside the window maps outside [0, dim).
__device__ __forceinline__ int win_x(int cx, int w0, int nc0)
{
	int l = cx - w0;
	if (l < 0) l += nc0;
	else if (l >= nc0) l -= nc0;
	return l;
}

__device__ __forceinline__ Window window_of(const int *bbox, const Geom &g, long long cellcap)
{
	Window w;
	long long n = 1;
	if (g.slab) {   // owned columns + halo on both sides, whole grid in y and z
		w.org[0] = g.col_lo - g.halo; w.dim[0] = (g.col_hi - g.col_lo) + 2 * g.halo;
		w.org[1] = 0; w.dim[1] = g.nc[1];
		w.org[2] = 0; w.dim[2] = g.nc[2];
		w.fd0 = w.dim[0] * g.xs;
		n = (long long)w.fd0 * w.dim[1] * w.dim[2];
		w.ncells = (n > cellcap) ? 0 : (int)n;
		return w;
	}
	for (int d = 0; d < 3; d++) {
		int lo = bbox[d], hi = bbox[3 + d];
		if (lo == INT_MAX) { lo = 0; hi = 0; }
		// one empty cell on each side; a particle in the first (last) cell of an axis moves to the last (first) one when it
		// wraps around the box, so an occupied box that touches a face takes the whole axis
		if (lo == 0 || hi == g.nc[d] - 1) { lo = 0; hi = g.nc[d] - 1; }
		lo = max(lo - 1, 0);
		hi = min(hi + 1, g.nc[d] - 1);
		w.org[d] = lo;
		w.dim[d] = hi - lo + 1;
		n *= w.dim[d];
	}
	w.fd0 = w.dim[0] * g.xs;
	n *= g.xs;
	w.ncells = (n > cellcap) ? 0 : (int)n;   // over capacity: flagged by k_scan3, nothing is binned
	return w;
}

// First pass of the counting sort FUSED into the kernel that moved the particle: key of p (k_bin's: window-local cell, x slice)
// under the window b.win, one warp-aggregated atomic per distinct key.  b.win was published by the previous build as the
// occupied box of ITS positions plus one cell on every side (k_scan, win_next), so a particle can only fall outside it by
// moving more than a whole cell beyond everything that was occupied in one step.  That is legal (the reference follows any
// motion inside the box): the window is then marked dirty and the build that follows redoes the histogram under the
// window of the new occupied box (k_scan's fallback).  Returns the key (-1: not binned).  Must be called by all 32 lanes.
__device__ __forceinline__ int bin_particle(const Particle &p, bool live, const Geom &g, const BinArgs &b, int *errflag)
{
	int local = -1;
	if (live) {
		int cx, cy, cz;
		unpack_cell(p.cell, cx, cy, cz);
		// (a slab's window is fixed -- its columns + halo, wrapping across the periodic seam: a particle outside it has moved
		// further than the halo in one step, which is an error there, not a reason to re-bin)
		const int lx = g.slab ? win_x(cx, b.win[WIN_ORG], g.nc[0]) : cx - b.win[WIN_ORG], ly = cy - b.win[WIN_ORG + 1], lz = cz - b.win[WIN_ORG + 2];
		const int d0 = b.win[WIN_DIM], d1 = b.win[WIN_DIM + 1], d2 = b.win[WIN_DIM + 2];
		if (b.win[WIN_NCELLS] == 0 || lx < 0 || lx >= d0 || ly < 0 || ly >= d1 || lz < 0 || lz >= d2) {
			if (g.slab) atomicOr(errflag, ERR_SLAB_MIGRATION);
			else b.win[WIN_DIRTY] = 1;
		} else {
			int sub = 0;
			if (g.xs > 1) sub = min(max((int)((p.x - (double)cx * g.cs[0]) * g.finv), 0), g.xs - 1);
			local = (lx * g.xs + sub) + b.win[WIN_FD0] * (ly + d1 * lz);
		}
	}
	const unsigned active = __ballot_sync(0xffffffffu, local >= 0);
	if (local >= 0) {
		const unsigned peers = __match_any_sync(active, local);
		if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(b.count + local, __popc(peers));
	}
	return local;
}

// zero the histogram a tagging pass filled when no build is going to consume it (the particles are replaced or rescaled first)
__global__ void __launch_bounds__(256) k_clear_count(const int *win, int *count)
{
	const int n = win[WIN_NCELLS];
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) count[i] = 0;
}

// pos16[]: the 8-byte phase-1 candidate record of k_pair_force2.  Coordinates are 16-bit fixed point relative to the
// window's origin with ONE resolution for the three axes (so r^2 needs no per-axis scaling): the longest window edge
// maps to just under 65535 steps.  A coordinate inside the window is off by at most 0.505 steps (round to nearest +
// the FP32 rounding of 1/res), a difference of two by 1.01, a difference vector by 1.75 steps (sqrt(3) * 1.01).
// The fourth half-word is the particle's class cutoff^2 in steps^2, ((R + 1.75 res) / res)^2 rounded UP to a bf16:
// a pair the exact FP64 test accepts always passes the 16-bit test, the rest are dropped again in phase 2.
__device__ __forceinline__ float quant_res(const Window &w, const Geom &g)
{
	double ext = 0;
	for (int d = 0; d < 3; d++) ext = fmax(ext, (double)w.dim[d] * g.cs[d]);
	return __double2float_ru(ext * (1.0 + 1e-6) / 65535.0);
}

__device__ __forceinline__ uint2 quantize16(const Particle &p, const Geom &g, const int *win, float res, float inv_res, float radius)
{
	unsigned q[3];
	const double c[3] = {p.x, p.y, p.z};
#pragma unroll
	for (int d = 0; d < 3; d++) {
		double rel = (c[d] - (double)win[WIN_ORG + d] * g.cs[d]) * (double)inv_res;
		// outside the window only for the ghost columns a slab keeps across the periodic seam; those are never looked
		// at through this record (their pairs go through the periodic-image path, absolute FP32 coordinates)
		q[d] = (unsigned)min(max(__double2int_rn(rel), 0), 65535);
	}
	unsigned thr = 0xFF80u;   // -inf: interacts with nothing
	if (radius >= 0.f) {
		float t = __fmaf_rn(radius, inv_res, 1.75f);
		t = (t * t) * 1.000001f;
		thr = (__float_as_uint(t) + 0xFFFFu) >> 16;
	}
	(void)res;
	return make_uint2(q[0] | (q[1] << 16), q[2] | (thr << 16));
}

// pass 1 of the counting sort: histogram over the window.  Warp-aggregated: lanes sharing a cell elect a leader
// that issues one atomicAdd for the group.
__global__ void __launch_bounds__(TPB) k_bin(Cnt cnt, const Particle *pos, Geom g, const int *bbox, long long cellcap, int *count,
                                             int *cellOfSlot, int *errflag, const int *__restrict__ gid, int *slot_of)
{
	pdl_prologue();
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	Window w = window_of(bbox, g, cellcap);
	int local = -1;
	if (s < cnt.get() && w.ncells > 0) {
		const Particle p = load_particle(pos + s);
		unsigned c = p.cell;
		// slab mode: a ghost of the previous step.  Its entry of the index table is given back here -- the exchange that
		// dropped it may have run inside the step seam, where other threads were still looking partners up through the
		// table; if the particle has just been received again, k_reorder (later in this build) re-enters it.
		if (c == CELL_DEAD && slot_of) slot_of[gid[s] & GID_MASK] = -1;
		if (c != CELL_DEAD) {
			int cx, cy, cz;
			unpack_cell(c, cx, cy, cz);
			int lx = win_x(cx, w.org[0], g.nc[0]);
			// x slice inside the reference cell (sort key only; clamped, so that slice / xs is always the reference's cell)
			int sub = 0;
			if (g.xs > 1) sub = min(max((int)((p.x - (double)cx * g.cs[0]) * g.finv), 0), g.xs - 1);
			if (lx >= w.dim[0]) atomicOr(errflag, ERR_SLAB_MIGRATION);   // cannot happen on a single GPU (window = bbox)
			else local = (lx * g.xs + sub) + w.fd0 * ((cy - w.org[1]) + w.dim[1] * (cz - w.org[2]));
		}
		cellOfSlot[s] = local;
	}
	unsigned active = __ballot_sync(0xffffffffu, local >= 0);
	if (local >= 0) {
		unsigned peers = __match_any_sync(active, local);
		int leader = __ffs(peers) - 1;
		if ((int)(threadIdx.x & 31) == leader) atomicAdd(count + local, __popc(peers));
	}
}

// start the occupied-cell extremes from scratch (after set_particles / a box move that changed the grid)
__global__ void k_arm_bbox(int *bbox)
{
	if (threadIdx.x < 3) { bbox[threadIdx.x] = INT_MAX; bbox[3 + threadIdx.x] = INT_MIN; }
}

// Exclusive scan of count[0..ncells) in ONE kernel (ncells lives on the device; the grid is fixed and persistent, all
// blocks resident): chunk totals, a scan of the totals by the block that arrives last, then the chunks themselves (see the
// body).  Measured first: decoupled look-back over tiles dealt round-robin -- with one wave of tiles every predecessor is
// still an aggregate, so tile t polls all t states: 16 k polls on a dozen cache lines of one L2 slice, 16 us.  Here a
// block polls one word.  Prefix word: {epoch : 30 | flag : 2 | value : 32}; the epoch (the build counter) makes the words of
// earlier builds read as "not yet", so nothing is ever cleared.
// Block 0 publishes the window of this build in win[] (the kernels that follow read it there) and, in win_next[], the
// window the NEXT tagging pass may bin into (see bin_particle): the occupied box of the positions of this build + 1 cell.
// prebinned: the histogram was filled by the tagging pass itself under the window already in win[]; else (after
// set_particles / a box move / in slab mode) k_bin filled it under window_of(bbox), which is published here.
// nlive: slab mode only -- the number of live local particles after this build
constexpr int SCAN_TILE = 8 * SCAN_TPB;
constexpr unsigned long long SCAN_AGG = 1ull, SCAN_PREFIX = 2ull;

__device__ __forceinline__ void publish_window(int *win, const Window &wd, const Geom &g)
{
	for (int d = 0; d < 3; d++) { win[WIN_ORG + d] = wd.org[d]; win[WIN_DIM + d] = wd.dim[d]; }
	win[WIN_NCELLS] = wd.ncells;
	win[WIN_FD0] = wd.fd0;
	const float res = quant_res(wd, g);
	win[WIN_RES] = __float_as_int(res);
	win[WIN_INVRES] = __float_as_int(1.0f / res);
	win[WIN_DIRTY] = 0;
}

// barrier over a grid whose blocks are all resident (k_scan's rarely taken fallback); *ctr only ever grows
__device__ __forceinline__ void grid_barrier(unsigned *ctr, int *errflag)
{
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		const unsigned ticket = atomicAdd(ctr, 1u);
		const unsigned target = (ticket / gridDim.x + 1u) * gridDim.x;
		const long long t0 = clock64();
		while ((int)(atomicAdd(ctr, 0u) - target) < 0)
			if (clock64() - t0 > 20000000000ll) { atomicOr(errflag, ERR_WINDOW_CAP); break; }   // ~10 s: never hang the device
		__threadfence();
	}
	__syncthreads();
}

__global__ void __launch_bounds__(SCAN_TPB) k_scan(int *count, const int *bbox, Geom g, long long cellcap, int *win, int *win_next, int prebinned,
                                                   int *start, int *cursor, int N, int *nlive, int *errflag,
                                                   unsigned long long *state, unsigned epoch, const Particle *pos, int *cellOfSlot,
                                                   unsigned *barrier)
{
	pdl_prologue();
	SMD_TL(0);
	__shared__ int sh[SCAN_TPB / 32];
	__shared__ int s_excl;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	int ncells;
	if (prebinned && win[WIN_DIRTY]) {
		// A particle left the window the tagging pass binned into (bin_particle): redo the histogram under the window of the
		// new occupied box.  The flag was final before this grid started, so every block takes this path together; the grid is
		// persistent (all blocks resident), which makes the two barriers safe.
		const int oldn = win[WIN_NCELLS];
		const int gtid = blockIdx.x * SCAN_TPB + tid, gsz = gridDim.x * SCAN_TPB;
		for (int i = gtid; i < oldn; i += gsz) count[i] = 0;
		grid_barrier(barrier, errflag);
		const Window wd = window_of(bbox, g, cellcap);
		if (blockIdx.x == 0 && tid == 0) publish_window(win, wd, g);
		for (int s0 = (gtid & ~31); s0 < N; s0 += gsz) {   // (whole warps: the aggregation votes)
			const int s = s0 + lane;
			int local = -1;
			if (s < N && wd.ncells > 0) {
				const Particle p = load_particle(pos + s);
				int cx, cy, cz;
				unpack_cell(p.cell, cx, cy, cz);
				int sub = 0;
				if (g.xs > 1) sub = min(max((int)((p.x - (double)cx * g.cs[0]) * g.finv), 0), g.xs - 1);
				local = ((cx - wd.org[0]) * g.xs + sub) + wd.fd0 * ((cy - wd.org[1]) + wd.dim[1] * (cz - wd.org[2]));
				cellOfSlot[s] = local;
			}
			const unsigned active = __ballot_sync(0xffffffffu, local >= 0);
			if (local >= 0) {
				const unsigned peers = __match_any_sync(active, local);
				if (lane == __ffs(peers) - 1) atomicAdd(count + local, __popc(peers));
			}
		}
		grid_barrier(barrier, errflag);
		ncells = wd.ncells;
	} else if (prebinned) ncells = win[WIN_NCELLS];
	else {
		const Window wd = window_of(bbox, g, cellcap);
		ncells = wd.ncells;
		if (blockIdx.x == 0 && tid == 0) publish_window(win, wd, g);
	}
	if (blockIdx.x == 0 && tid == 0) {
		if (ncells == 0) atomicOr(errflag, ERR_WINDOW_CAP);
		if (win_next) publish_window(win_next, window_of(bbox, g, cellcap), g);
	}
	// ---- every block owns a contiguous chunk of tiles.  Pass 1: the chunk's total.  The block that arrives last (a ticket
	// counter that only ever grows: every build all blocks take exactly one ticket) scans the gridDim.x totals and
	// publishes one exclusive prefix per block, stamped with the build's epoch; a block polls nothing but its own word.
	// Pass 2: the chunk again (L2-hot), scanned tile by tile from that prefix, written out, histogram zeroed.
	const int T = (ncells + SCAN_TILE - 1) / SCAN_TILE;
	const int tpb = (T + (int)gridDim.x - 1) / (int)gridDim.x;
	const int t0 = min((int)blockIdx.x * tpb, T), t1 = min(t0 + tpb, T);
	const unsigned long long ep = (unsigned long long)(epoch & 0x3fffffffu) << 34;
	auto load8 = [&](int i, int (&v)[8]) {
		if (i + 7 < ncells) {
			const int4 a = *reinterpret_cast<const int4 *>(count + i), b = *reinterpret_cast<const int4 *>(count + i + 4);
			v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
		} else {
#pragma unroll
			for (int k = 0; k < 8; k++) v[k] = (i + k < ncells) ? count[i + k] : 0;
		}
	};
	auto block_total = [&](int x) {   // sum over the block, returned to every thread
		x = __reduce_add_sync(0xffffffffu, x);
		__syncthreads();
		if (lane == 0) sh[w] = x;
		__syncthreads();
		int tot = 0;
#pragma unroll
		for (int k = 0; k < SCAN_TPB / 32; k++) tot += sh[k];
		return tot;
	};
	int mine = 0;
	for (int t = t0; t < t1; t++) {
		int v[8];
		load8(t * SCAN_TILE + 8 * tid, v);
#pragma unroll
		for (int k = 0; k < 8; k++) mine += v[k];
	}
	const int chunk_total = block_total(mine);
	__shared__ int s_last;
	int *agg = reinterpret_cast<int *>(state + gridDim.x);   // [gridDim.x] chunk totals of this build, behind the prefix words
	if (tid == 0) {
		agg[blockIdx.x] = chunk_total;
		__threadfence();
		const unsigned ticket = atomicAdd(barrier + 1, 1u);
		s_last = (ticket % gridDim.x) == gridDim.x - 1;
	}
	__syncthreads();
	if (s_last) {   // (uniform per block) everybody's total is in: one block scans them
		__threadfence();
		int carry = 0;
		for (int b0 = 0; b0 < (int)gridDim.x; b0 += SCAN_TPB) {
			const int b = b0 + tid;
			const int x = b < (int)gridDim.x ? *reinterpret_cast<volatile int *>(agg + b) : 0;
			int inc = x;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
			__syncthreads();
			if (lane == 31) sh[w] = inc;
			__syncthreads();
			int woff = 0, total = 0;
#pragma unroll
			for (int k = 0; k < SCAN_TPB / 32; k++) { const int y = sh[k]; if (k < w) woff += y; total += y; }
			if (b < (int)gridDim.x) {
				const unsigned long long word = ep | (SCAN_PREFIX << 32) | (unsigned)(carry + woff + inc - x);
				asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(state + b), "l"(word) : "memory");
			}
			carry += total;
		}
	}
	if (tid == 0) {   // this block's prefix (a word carries everything the reader needs: relaxed accesses)
		unsigned long long word;
		do {
			asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(word) : "l"(state + blockIdx.x) : "memory");
			if ((word >> 34) != (ep >> 34)) __nanosleep(200);
		} while ((word >> 34) != (ep >> 34));
		s_excl = (int)(unsigned)(word & 0xffffffffull);
	}
	__syncthreads();
	int carry = s_excl;
	for (int t = t0; t < t1; t++) {
		const int i = t * SCAN_TILE + 8 * tid;
		int v[8];
		load8(i, v);
		int sum8 = 0;
#pragma unroll
		for (int k = 0; k < 8; k++) sum8 += v[k];
		int inc = sum8;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
		__syncthreads();
		if (lane == 31) sh[w] = inc;
		__syncthreads();
		int woff = 0, total = 0;
#pragma unroll
		for (int k = 0; k < SCAN_TPB / 32; k++) { const int x = sh[k]; if (k < w) woff += x; total += x; }
		int ex = carry + woff + inc - sum8;
		int e[8];
#pragma unroll
		for (int k = 0; k < 8; k++) { e[k] = ex; ex += v[k]; }
		if (i + 7 < ncells) {
			const int4 a = make_int4(e[0], e[1], e[2], e[3]), b = make_int4(e[4], e[5], e[6], e[7]), z = make_int4(0, 0, 0, 0);
			*reinterpret_cast<int4 *>(start + i) = a; *reinterpret_cast<int4 *>(start + i + 4) = b;
			*reinterpret_cast<int4 *>(cursor + i) = a; *reinterpret_cast<int4 *>(cursor + i + 4) = b;
			*reinterpret_cast<int4 *>(count + i) = z; *reinterpret_cast<int4 *>(count + i + 4) = z;
		} else {
#pragma unroll
			for (int k = 0; k < 8; k++)
				if (i + k < ncells) { start[i + k] = e[k]; cursor[i + k] = e[k]; count[i + k] = 0; }
		}
		carry += total;
		if (t == T - 1 && tid == 0) {   // the carry behind the last tile = the number of binned particles
			const int all = nlive ? carry : N;
			start[ncells] = all;
			if (nlive) *nlive = all;
		}
	}
	if (T == 0 && blockIdx.x == 0 && tid == 0) { start[0] = nlive ? 0 : N; if (nlive) *nlive = 0; }
}

// pass 2: claim a position inside the cell's range (arbitrary order, fixed by k_reorder)
__global__ void __launch_bounds__(TPB) k_place(Cnt cnt, const int *cellOfSlot, int *cursor, int2 *order, const int *__restrict__ gid,
                                               int *giveback)
{
	pdl_prologue();
	SMD_TL(1);
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= cnt.get()) return;
	int c = cellOfSlot[s];
	if (c < 0) {
		// slab mode, histogram filled by the seam / the unpack (no k_bin in this build): a record that was not binned is a
		// ghost the exchange dropped; its entry of the index table is given back here (see k_bin), before k_reorder re-enters
		// the particle if it has just been received again
		if (giveback) giveback[gid[s] & GID_MASK] = -1;
		return;
	}
	int q = atomicAdd(cursor + c, 1);
	order[q] = make_int2(s, gid[s] & GID_MASK);   // the original index travels along: k_reorder ranks a cell without a second gather
}

// pass 3: final slot = cell start + rank by DESCENDING original index, which is exactly the order of the
// reference's head-inserted linked list (cellOpt.h:572-585) and makes the sort deterministic; move the records.
__global__ void __launch_bounds__(TPB) k_reorder(Cnt cnt, int cap, const int2 *__restrict__ order, const int *cellOfSlot, const int *start,
                                                 const Particle *pos_in, Particle *pos_out, const double *vel_in, double *vel_out,
                                                 const double *unw_in, double *unw_out, const double *acc_in, double *acc_out,
                                                 const int *gid_in, int *gid_out, int *slot_of, float4 *pos32_out,
                                                 const float *__restrict__ acut, int *bbox, int rearm, uint2 *pos16_out,
                                                 const float *__restrict__ arad, const int *__restrict__ win, Geom geo)
{
	pdl_prologue();
	SMD_TL(2);
	int q = blockIdx.x * blockDim.x + threadIdx.x;
	// rearm: every few hundred builds the occupied-cell extremes start from scratch, to follow a drifting object (nobody
	// reads them between the scan and the next tagging pass)
	if (rearm && q < 3) { bbox[q] = INT_MAX; bbox[3 + q] = INT_MIN; }
	if (q >= cnt.get()) return;
	const int2 me = order[q];
	int s = me.x;
	int c = cellOfSlot[s];
	int b = start[c], e = start[c + 1];
	int g = gid_in[s];
	int rank = 0;
	for (int k = b; k < e; k++) rank += (order[k].y > me.y);
	int t = b + rank;
	Particle p = load_particle(pos_in + s);
	store_particle(pos_out + t, p);
	pos32_out[t] = make_float4((float)p.x, (float)p.y, (float)p.z, acut[p.type]);   // FP32 mirror: periodic-image path of k_pair_force2
	pos16_out[t] = quantize16(p, geo, win, __int_as_float(win[WIN_RES]), __int_as_float(win[WIN_INVRES]), arad[p.type]);   // its phase-1 candidates
	vel_out[t] = vel_in[s]; vel_out[cap + t] = vel_in[cap + s]; vel_out[2 * cap + t] = vel_in[2 * cap + s];
	if (unw_in) { unw_out[t] = unw_in[s]; unw_out[cap + t] = unw_in[cap + s]; unw_out[2 * cap + t] = unw_in[2 * cap + s]; }
	if (acc_in) { acc_out[t] = acc_in[s]; acc_out[cap + t] = acc_in[cap + s]; acc_out[2 * cap + t] = acc_in[2 * cap + s]; }
	gid_out[t] = g;
	slot_of[g & GID_MASK] = t;
}

// ------------------------------------------------------------------------------------------------ pair engine
// MD.h:795-848 Force<T>, MD.h:895-930 Potential<T>

__device__ __forceinline__ int pair_branch(double dr, const double *fc)
{
	// int(dr / fC[cindex]) * 3 (MD.h:822).  For dr < 2 rm the quotient's integer part is decided exactly by
	// comparisons (IEEE division is monotone and dr < rm => fl(dr/rm) < 1, dr < 2rm => fl(dr/rm) < 2), so the
	// division only runs in the never-used region dr >= 2 rm.
	double rm = fc[0];
	if (dr < rm) return 0;
	if (dr < rm + rm) return 3;
	return ((int)(dr / rm)) * 3;
}

__device__ __forceinline__ double pair_force_mag(double dr2, int t1, int t2, int nT, const double *fC, int ntab)
{
	double dr = sqrt(dr2);
	int ci = 6 * (t1 * nT + t2);
	ci += pair_branch(dr, fC + ci);
	ci = min(ci, ntab - 3);   // the reference reads past its table for dr >= 3 rm in the last row; stay in bounds
	double m = fC[ci] - dr;
	return ((fC[ci + 1] - fC[ci + 2] * m) * m) / dr;
}

__device__ __forceinline__ double pair_potential_val(double dr2, int t1, int t2, int nT, const double *uC)
{
	double dr = sqrt(dr2);
	int ci = 6 * (t1 * nT + t2);
	double u;
	if (dr <= uC[ci]) {
		u = uC[ci] - dr;
		u = uC[ci + 1] * u * u + uC[ci + 2];
	} else {
		u = uC[ci + 3] - dr;
		u = u * u * (uC[ci + 4] - u * uC[ci + 5]);
	}
	return u;
}

enum PairMode { PAIR_FORCE = 0, PAIR_POTENTIAL = 1, PAIR_DPOTENTIAL = 2, PAIR_COUNT = 3 };

// One thread per particle (slot).  The reference walks a half stencil (13 forward cells, cellOpt.h:715) with
// Newton's third law under per-cell locks; here every particle gathers its own force from the full 27-cell stencil:
// no atomics, deterministic, and each pair term is BIT-IDENTICAL to the reference's because
//   * d = p1 - p2 is exactly antisymmetric, so evaluating a pair from either end gives the same |d|^2 and magnitude;
//   * for a neighbour cell reached across the periodic boundary the reference adds +-L to the NEIGHBOUR particle
//     of the HOME cell before subtracting (cellOpt.h:811-821,846-848).  We reproduce that orientation: a forward
//     offset shifts the neighbour, a backward offset means the neighbour's cell is the home cell and WE get shifted;
//   * the constant row is picked in the reference's orientation (type of the home / later-loaded particle first).
// PAIR_POTENTIAL / PAIR_DPOTENTIAL visit every pair once (forward cells + lower index first in the own cell) and
// reduce per block.  tab = fC for force, uC otherwise (staged in shared memory).
template <int MODE>
__global__ void __launch_bounds__(TPB) k_pair(Cnt cnt, int cap, const Particle *__restrict__ pos, const int *__restrict__ gid,
                                              const int *__restrict__ start, const int *__restrict__ win, Geom g, int nT,
                                              const double *__restrict__ tab, double *__restrict__ acc, double *__restrict__ partials,
                                              int *__restrict__ icount, double sx, double sy, double sz)
{
	extern __shared__ double s_tab[];
	int ntab = 6 * nT * nT;
	for (int k = threadIdx.x; k < ntab; k += blockDim.x) s_tab[k] = tab[k];
	__syncthreads();

	int i = blockIdx.x * blockDim.x + threadIdx.x;
	double ax = 0, ay = 0, az = 0, usum = 0;
	int cnt_in = 0;
	// slab mode: only owned particles gather / count; a pair of an owned and a ghost particle is counted by the rank
	// that owns the reference's home particle of the pair, so every pair is still visited exactly once globally
	if (i < cnt.get() && !(g.slab && (gid[i] & GID_GHOST))) {
		Particle pi = load_particle(pos + i);
		int cx, cy, cz;
		unpack_cell(pi.cell, cx, cy, cz);
		int w0 = win[WIN_ORG], w1 = win[WIN_ORG + 1], w2 = win[WIN_ORG + 2];
		int d0 = win[WIN_DIM], d1 = win[WIN_DIM + 1], d2 = win[WIN_DIM + 2];
		for (int oz = -1; oz <= 1; oz++) {
			int nz = cz + oz;
			double Sz = 0;
			if (nz < 0) { nz += g.nc[2]; Sz = -g.box[2]; }
			if (nz >= g.nc[2]) { nz -= g.nc[2]; Sz = g.box[2]; }
			int lz = nz - w2;
			if (lz < 0 || lz >= d2) continue;
			for (int oy = -1; oy <= 1; oy++) {
				int ny = cy + oy;
				double Sy = 0;
				if (ny < 0) { ny += g.nc[1]; Sy = -g.box[1]; }
				if (ny >= g.nc[1]) { ny -= g.nc[1]; Sy = g.box[1]; }
				int ly = ny - w1;
				if (ly < 0 || ly >= d1) continue;
				for (int ox = -1; ox <= 1; ox++) {
					int nx = cx + ox;
					double Sx = 0;
					if (nx < 0) { nx += g.nc[0]; Sx = -g.box[0]; }
					if (nx >= g.nc[0]) { nx -= g.nc[0]; Sx = g.box[0]; }
					int lx = win_x(nx, w0, g.nc[0]);
					if (lx >= d0) continue;
					bool self = (ox == 0 && oy == 0 && oz == 0);
					// forward offsets of cellOpt.h:715
					bool fwd = (oz == 1) || (oz == 0 && (ox == 1 || (ox == 0 && oy == 1)));
					if (MODE == PAIR_POTENTIAL || MODE == PAIR_DPOTENTIAL)
						if (!self && !fwd) continue;
					int c = lx * g.xs + win[WIN_FD0] * (ly + d1 * lz);   // the cell's x slices are consecutive table entries
					int jb = start[c], je = start[c + g.xs];
					bool shifted = (Sx != 0) || (Sy != 0) || (Sz != 0);
					// backward: the neighbour's cell is home and shifts US by the opposite image vector
					double bx = pi.x - Sx, by = pi.y - Sy, bz = pi.z - Sz;
					for (int j = jb; j < je; j++) {
						if (self && j == i) continue;
						Particle pj = load_particle(pos + j);
						double dx, dy, dz;
						if (!shifted || self) {
							dx = pi.x - pj.x; dy = pi.y - pj.y; dz = pi.z - pj.z;
						} else if (fwd) {
							dx = pi.x - (pj.x + Sx); dy = pi.y - (pj.y + Sy); dz = pi.z - (pj.z + Sz);
						} else {
							dx = -(pj.x - bx); dy = -(pj.y - by); dz = -(pj.z - bz);
						}
						double dr2 = dx * dx + dy * dy + dz * dz;
						if (MODE == PAIR_COUNT) {
							cnt_in += (dr2 < g.rc2);
							continue;
						}
						// am I the reference's p1 (home cell; in the own cell the later-loaded = lower original index,
						// i.e. the HIGHER slot because a cell's slots are sorted by descending original index)?
						bool home = self ? (i > j) : fwd;
						if (MODE == PAIR_POTENTIAL || MODE == PAIR_DPOTENTIAL)
							if (!home) continue;
						int t1 = home ? pi.type : pj.type, t2 = home ? pj.type : pi.type;
						if (MODE == PAIR_FORCE) {
							if (dr2 < g.rc2) {
								double m = pair_force_mag(dr2, t1, t2, nT, s_tab, ntab);
								ax += dx * m; ay += dy * m; az += dz * m;
							}
						} else if (MODE == PAIR_POTENTIAL) {
							if (dr2 < g.rc2) usum += pair_potential_val(dr2, t1, t2, nT, s_tab);
						} else {
							// cellOpt.h:1097-1107 / :1161-1171: old - new with both (shifted) positions scaled
							double uo = (dr2 < g.rc2) ? pair_potential_val(dr2, t1, t2, nT, s_tab) : 0.0;
							double qx = pj.x, qy = pj.y, qz = pj.z;
							if (shifted && !self) { qx += Sx; qy += Sy; qz += Sz; }
							double ex = pi.x * sx - qx * sx, ey = pi.y * sy - qy * sy, ez = pi.z * sz - qz * sz;
							double er2 = ex * ex + ey * ey + ez * ez;
							double un = (er2 < g.rc2) ? pair_potential_val(er2, t1, t2, nT, s_tab) : 0.0;
							usum += (uo - un);
						}
					}
				}
			}
		}
		if (MODE == PAIR_FORCE) {
			acc[i] += ax; acc[cap + i] += ay; acc[2 * cap + i] += az;
		}
		if (MODE == PAIR_COUNT) icount[i] = cnt_in;
	}
	if (MODE == PAIR_POTENTIAL || MODE == PAIR_DPOTENTIAL) {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
	if (MODE == PAIR_COUNT) {
		double c = block_sum((double)cnt_in);
		if (threadIdx.x == 0) partials[blockIdx.x] = c;
	}
}

// ------------------------------------------------------------------------------------------------ bonded terms
// minimum image per component with strict compares, system.h:1809-1817
__device__ __forceinline__ void min_image(double &dx, double &dy, double &dz, const Geom &g)
{
	if (dx > g.box[0] / 2.0) dx -= g.box[0];
	if (dx < -g.box[0] / 2.0) dx += g.box[0];
	if (dy > g.box[1] / 2.0) dy -= g.box[1];
	if (dy < -g.box[1] / 2.0) dy += g.box[1];
	if (dz > g.box[2] / 2.0) dz -= g.box[2];
	if (dz < -g.box[2] / 2.0) dz += g.box[2];
}

struct V3 { double x, y, z; };

__device__ __forceinline__ V3 diff_mi(const Particle &a, const Particle &b, const Geom &g)
{
	V3 d = {a.x - b.x, a.y - b.y, a.z - b.z};
	min_image(d.x, d.y, d.z, g);
	return d;
}

// Square root and reciprocal for the bonded force terms.  The reference divides by r = sqrt(d.d) fourteen times per
// CHAIN triplet (MD.h:384-403, :683-723); IEEE sqrt / division sequences made the fused step kernel FP64-latency bound.
// Here one MUFU.RSQ64H seed per distinct length gives r = sqrt(x) (correctly rounded, as in the pair kernel) and
// 1 / r to full precision, and every quotient n / r is one multiply plus an exact-residual correction
// (q = n * inv; q += fma(-r, q, n) * inv): correctly rounded except for vanishingly rare near-ties (then one ulp off).
struct SqrtRcp { double r, inv; };

__device__ __forceinline__ SqrtRcp sqrt_rcp(double x)
{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	double e = __fma_rn(-(x * y), y, 1.0);
	y = __fma_rn(0.5 * y, e, y);                              // 1/sqrt(x) to ~2^-43
	double r = x * y;
	r = __fma_rn(__fma_rn(-r, r, x), 0.5 * y, r);             // sqrt(x)
	y = __fma_rn(__fma_rn(-r, y, 1.0), y, y);                 // 1/r, refined against the rounded r
	y = __fma_rn(__fma_rn(-r, y, 1.0), y, y);
	SqrtRcp o = {r, y};
	if (!(x > 0.0) || x == INFINITY) { o.r = sqrt(x); o.inv = 1.0 / o.r; }   // coincident particles / NaN: IEEE root and reciprocal (the run is lost either way)
	return o;
}

__device__ __forceinline__ double div_cr(double n, const SqrtRcp &d)
{
	double q = n * d.inv;
	return __fma_rn(__fma_rn(-d.r, q, n), d.inv, q);
}

// MD.h:384-403 harmonicF: returns f with a1 += f, a2 -= f
__device__ __forceinline__ V3 harmonic_f(V3 d, double r0, double k)
{
	SqrtRcp dr = sqrt_rcp(d.x * d.x + d.y * d.y + d.z * d.z);
	double m = dr.r - r0;
	m = div_cr(-m * k, dr);
	V3 f = {d.x * m, d.y * m, d.z * m};
	return f;
}

// MD.h:422-431 harmonicP
__device__ __forceinline__ double harmonic_p(V3 d, double r0, double k)
{
	double dr = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
	double u = dr - r0;
	return 0.5 * k * u * u;
}

// MD.h:683-723 bendF: a1 += fa ; a2 += (fb - fa) ; a3 -= fb
__device__ __forceinline__ void bend_f(V3 da, V3 db, double c0, double k, V3 &fa, V3 &fb)
{
	const SqrtRcp dra = sqrt_rcp(da.x * da.x + da.y * da.y + da.z * da.z);
	const SqrtRcp drb = sqrt_rcp(db.x * db.x + db.y * db.y + db.z * db.z);
	da.x = div_cr(da.x, dra); da.y = div_cr(da.y, dra); da.z = div_cr(da.z, dra);
	db.x = div_cr(db.x, drb); db.y = div_cr(db.y, drb); db.z = div_cr(db.z, drb);
	double ct = (da.x * db.x) + (da.y * db.y) + (da.z * db.z);
	double m = c0 - ct;
	m *= k;
	fa.x = div_cr(m * (db.x - (da.x * ct)), dra); fb.x = div_cr(m * (da.x - (db.x * ct)), drb);
	fa.y = div_cr(m * (db.y - (da.y * ct)), dra); fb.y = div_cr(m * (da.y - (db.y * ct)), drb);
	fa.z = div_cr(m * (db.z - (da.z * ct)), dra); fb.z = div_cr(m * (da.z - (db.z * ct)), drb);
}

// MD.h:769-789 bendP
__device__ __forceinline__ double bend_p(V3 da, V3 db, double c0, double k)
{
	double dra = sqrt(da.x * da.x + da.y * da.y + da.z * da.z);
	double drb = sqrt(db.x * db.x + db.y * db.y + db.z * db.z);
	da.x /= dra; da.y /= dra; da.z /= dra;
	db.x /= drb; db.y /= drb; db.z /= drb;
	double ct = (da.x * db.x) + (da.y * db.y) + (da.z * db.z);
	double u = c0 - ct;
	return k * u * u * 0.5;
}

__device__ __forceinline__ V3 scaled(V3 d, double sx, double sy, double sz)
{
	V3 r = {d.x * sx, d.y * sy, d.z * sz};
	return r;
}

// CHAIN blocks, one thread per chain, same triplet order as Blob::doChainForce (system.h:1782-1866) /
// doChainPotential (:2487-2568) / doChainDPotential (:3280-3384).  Chains are disjoint, so the force variant
// updates acc[] without atomics; each particle's chain terms are summed in the reference's order.
// MODE 0 force, 1 potential, 2 dPotential.
// slot of the particle with global index id on this rank, or -1 (slab mode: owned or ghost copy; see k_chain_slab)
__device__ __forceinline__ int slab_find(const int *__restrict__ slot_of, const int *__restrict__ gid, int N, int id, bool &owned)
{
	int t = slot_of[id];
	owned = false;
	if (t < 0 || t >= N) return -1;
	int gg = gid[t];
	if ((gg & GID_MASK) != id) return -1;   // stale entry of a particle that left this rank
	owned = !(gg & GID_GHOST);
	return t;
}

// ------------------------------------------------------------------------------------------------ fused step seam
// Between the pair force of step i and the cell build of step i+1 the reference runs, per particle, Blob::doChainForce
// (system.h:1782-1866), Verlet::second (verlet.h:463-477) and -- next iteration -- Verlet::first (verlet.h:288-356).
// For systems whose only molecules are CHAIN blocks (C1, C2, C5) k_chain_kick does all of it in ONE pass with one
// thread per particle: the particle gathers its own chain terms (the one to three triplets it belongs to, evaluated
// and accumulated in the reference's order, so the sum a_chain is bit-identical to the per-chain kernel's), adds them
// to the pair + Langevin acceleration, applies the two half kicks one after the other (two roundings, as two kernels
// would), drifts, wraps and tags the new cell.  New positions go to the other position buffer: neighbours still read
// the old ones for their chain terms.  Three latency-bound passes (43 us on C2) become one.
// LAST = true ends a batch of steps instead: chain terms + Verlet::second only, a[] is stored for whoever comes next.
constexpr int MAX_FUSED_CHAINS = 4;
struct ChainSet { int n; ChainBlock b[MAX_FUSED_CHAINS]; };

// the chain terms of the particle with global index gi (role by role, system.h:1798-1863)
__device__ __forceinline__ bool chain_gather(int gi, const Particle &me, int N, const Particle *__restrict__ pos, const int *__restrict__ gid,
                                             const int *__restrict__ slot_of, const Geom &g, const ChainSet &cs, V3 &A)
{
	A.x = A.y = A.z = 0;
	for (int b = 0; b < cs.n; b++) {
		const ChainBlock &cb = cs.b[b];
		int rel = gi - cb.start;
		if (rel < 0 || rel >= cb.nChains * cb.len) continue;
		int k = rel / cb.len, l = rel - k * cb.len, base = cb.start + k * cb.len;
		// members l-2 .. l+2 that exist
		Particle q[5];
		bool have[5];
#pragma unroll
		for (int d = 0; d < 5; d++) {
			int m = l + d - 2;
			have[d] = false;
			if (d == 2) { q[d] = me; have[d] = true; continue; }
			if (m < 0 || m >= cb.len) continue;
			// which triplets need member m: only those that contain l
			// slab mode: slot_of[] is exact after every build (the exchange clears the entries of dropped ghosts), so
			// a non-negative entry is a local particle, owned or ghost
			const int t = slot_of[base + m];
			if (t >= 0) { q[d] = load_particle(pos + t); have[d] = true; }
		}
		// The particle's triplets in ascending order = role 2 (triplet l-2), then role 1, then role 0: the reference's
		// order.  The loop runs over "my j-th triplet", not over roles, so that the lanes of a warp -- particles at
		// different positions of their chains -- evaluate their square roots and divisions together; only the few
		// additions that differ between the roles diverge.
		bool ok = true;
		const int tmin = max(l - 2, 0), tmax = min(l, cb.len - 3);
#pragma unroll
		for (int j = 0; j < 3; j++) {
			const int t = tmin + j;
			if (t > tmax) break;
			const int r = l - t;                // my position inside the triplet
			const bool tail = (t == cb.len - 3);
			const bool r2 = (r == 2), r1 = (r == 1);
			const bool h0 = r2 ? have[0] : (r1 ? have[1] : have[2]);
			const bool h1 = r2 ? have[1] : (r1 ? have[2] : have[3]);
			const bool h2 = r2 ? have[2] : (r1 ? have[3] : have[4]);
			if (!(h0 && h1 && h2)) { ok = false; continue; }
			const Particle &P0 = r2 ? q[0] : (r1 ? q[1] : q[2]);
			const Particle &P1 = r2 ? q[1] : (r1 ? q[2] : q[3]);
			const Particle &P2 = r2 ? q[2] : (r1 ? q[3] : q[4]);
			V3 da = diff_mi(P0, P1, g), db = diff_mi(P1, P2, g);
			V3 fa, fb;
			bend_f(da, db, cb.c[2], cb.c[3], fa, fb);
			V3 f = harmonic_f(da, cb.c[0], cb.c[1]);
			V3 f2 = {0, 0, 0};
			if (tail) f2 = harmonic_f(db, cb.c[0], cb.c[1]);
			if (r2) {
				if (tail) { A.x -= f2.x; A.y -= f2.y; A.z -= f2.z; }
				A.x -= fb.x; A.y -= fb.y; A.z -= fb.z;
			} else if (r1) {
				A.x -= f.x; A.y -= f.y; A.z -= f.z;
				if (tail) { A.x += f2.x; A.y += f2.y; A.z += f2.z; }
				A.x += (fb.x - fa.x); A.y += (fb.y - fa.y); A.z += (fb.z - fa.z);
			} else {
				A.x += f.x; A.y += f.y; A.z += f.z;
				A.x += fa.x; A.y += fa.y; A.z += fa.z;
			}
		}
		return ok;
	}
	return true;
}

// ------------------------------------------------------------------------------------------------ pair force, two-phase
// The production force kernel.  Same pairs, same per-pair arithmetic and the same orientation rules as k_pair above
// (so every pair term stays bit-identical to the reference's), reorganised around what the B200 is short of here:
// FP64 issue slots.  ncu on k_pair showed 11.4 of 32 lanes active per instruction: the sqrt / divide body ran under
// a 20 % hit-rate branch.
//   phase 1  (FP32 / INT / LSU pipes): each lane walks the 9 (y,z) rows of its particle's stencil -- the 3 cells of a
//            row are contiguous in the sorted order, so a row is one index range -- over a float4 mirror of the
//            positions and tests r^2 < rc^2 + margin in FP32.  Candidates that pass are appended to a lane-private
//            list in shared memory (column `lane` of a [CAP][32] array: conflict-free).
//   phase 2  (FP64 pipe): every lane drains its own list: reload the neighbour's FP64 record, repeat the exact FP64
//            test (bit-exact membership), evaluate the pair term, accumulate in registers.  Lanes now run the
//            expensive body together; only the spread of list lengths (63 +- 8) idles lanes.
// The margin covers the FP32 rounding of absolute coordinates (<= box * 2^-24 each), so phase 1 never drops a pair
// that the FP64 test accepts; false positives are rejected in phase 2.
// The Langevin term and the zeroing of a[] are folded into the epilogue when LANGEVIN is set:
//   a = (-gamma v + sigma (2u-1)) + sum_pairs,   exactly the order MD.cpp:357-413 produces.
constexpr int PAIR_TPB = 128;
#ifndef SMD_PAIR_CAP
#define SMD_PAIR_CAP 128
#endif
#ifndef SMD_PAIR_BLOCKS
#define SMD_PAIR_BLOCKS 4
#endif
constexpr int PAIR_CAP = SMD_PAIR_CAP;   // list entries per lane, 16 bit each: 4 warps x 128 x 32 x 2 B = 32 KiB per block
// k_pair_force2<.., SPLIT>: SPLIT = 1: 128 particles per block, one thread each.  SPLIT = 3: 32 particles per block, three
// threads each (one per z plane of the stencil) -- a shorter critical path for systems that do not fill the device.  (Nine
// threads each, one per (y,z) row, was built and measured: 58 against 25 us per launch on 15 000 particles -- the per-thread
// set-up and the block-wide hand-over are repeated nine times; profiles/r02b_pair3_ab.md.)
#ifndef SMD_PAIR3_BLOCKS
#define SMD_PAIR3_BLOCKS 4
#endif
#ifndef SMD_PAIR3_CAP
#define SMD_PAIR3_CAP 64
#endif
template <int SPLIT> struct PairCfg {
	static constexpr int NP = SPLIT == 1 ? 128 : 32;                                   // particles per block
	static constexpr int BT = NP * SPLIT;                                              // threads per block
	static constexpr int CAP = SPLIT == 1 ? SMD_PAIR_CAP : SMD_PAIR3_CAP;   // list entries per thread
	static constexpr int BLOCKS = SPLIT == 1 ? SMD_PAIR_BLOCKS : SMD_PAIR3_BLOCKS;   // resident blocks per SM (register budget)
	static constexpr int NROW = 9 / SPLIT;                                             // stencil rows per thread
};
constexpr int PAIR_SPLIT_NP = 32;   // particles per block of the split engine (the unit of its completion words)
constexpr int PAIR_SEGBITS = 12;   // entry = (range index << 12) | offset inside the range

struct LangevinArgs { double gamma, sigma; uint64_t seed, step; const double *vel; const int *gid; const double *ext_noise; };

// the general pair term: any orientation, periodic images, ordered (possibly asymmetric) constant tables.
// Same arithmetic as k_pair<PAIR_FORCE>.  Off the fast path, so kept out of line.
struct D3 { double x, y, z; };

__device__ __noinline__ D3 pair_force_term(int i, Particle pi, int j, Particle pj, const Geom &g, int nT, const double *s_tab, int ntab)
{
	D3 f = {0.0, 0.0, 0.0};
	int cx, cy, cz, jx, jy, jz;
	unpack_cell(pi.cell, cx, cy, cz);
	unpack_cell(pj.cell, jx, jy, jz);
	int ox = jx - cx, oy = jy - cy, oz = jz - cz;
	double Sx = 0, Sy = 0, Sz = 0;
	// neighbour cell reached across the periodic boundary: cellOpt.h:811-821
	if (ox > 1) { ox -= g.nc[0]; Sx = -g.box[0]; } else if (ox < -1) { ox += g.nc[0]; Sx = g.box[0]; }
	if (oy > 1) { oy -= g.nc[1]; Sy = -g.box[1]; } else if (oy < -1) { oy += g.nc[1]; Sy = g.box[1]; }
	if (oz > 1) { oz -= g.nc[2]; Sz = -g.box[2]; } else if (oz < -1) { oz += g.nc[2]; Sz = g.box[2]; }
	bool self = (ox == 0 && oy == 0 && oz == 0);
	bool fwd = (oz == 1) || (oz == 0 && (ox == 1 || (ox == 0 && oy == 1)));
	bool shifted = (Sx != 0) || (Sy != 0) || (Sz != 0);
	double dx, dy, dz;
	if (!shifted || self) {
		dx = pi.x - pj.x; dy = pi.y - pj.y; dz = pi.z - pj.z;
	} else if (fwd) {
		dx = pi.x - (pj.x + Sx); dy = pi.y - (pj.y + Sy); dz = pi.z - (pj.z + Sz);
	} else {
		double bx = pi.x - Sx, by = pi.y - Sy, bz = pi.z - Sz;
		dx = -(pj.x - bx); dy = -(pj.y - by); dz = -(pj.z - bz);
	}
	double dr2 = dx * dx + dy * dy + dz * dz;
	if (dr2 < g.rc2) {
		bool home = self ? (i > j) : fwd;
		int t1 = home ? pi.type : pj.type, t2 = home ? pj.type : pi.type;
		double m = pair_force_mag(dr2, t1, t2, nT, s_tab, ntab);
		f.x = dx * m; f.y = dy * m; f.z = dz * m;
	}
	return f;
}

// energy twin of pair_force_term for the two-phase kernel's slow path (pairs seen through a periodic image): the term
// of the pair (i, j) if i is the reference's home particle of it, else 0 -- both ends call it, one of them counts.
// EMODE 1: U (cellOpt.h:928-1041), 2: U(d) - U(d o scale) (cellOpt.h:1043-1180).  Same arithmetic as k_pair.
template <int EMODE>
__device__ __noinline__ double pair_energy_term(int i, Particle pi, int j, Particle pj, const Geom &g, int nT, const double *uC,
                                                double sx, double sy, double sz, bool always)
{
	int cx, cy, cz, jx, jy, jz;
	unpack_cell(pi.cell, cx, cy, cz);
	unpack_cell(pj.cell, jx, jy, jz);
	int ox = jx - cx, oy = jy - cy, oz = jz - cz;
	double Sx = 0, Sy = 0, Sz = 0;
	if (ox > 1) { ox -= g.nc[0]; Sx = -g.box[0]; } else if (ox < -1) { ox += g.nc[0]; Sx = g.box[0]; }
	if (oy > 1) { oy -= g.nc[1]; Sy = -g.box[1]; } else if (oy < -1) { oy += g.nc[1]; Sy = g.box[1]; }
	if (oz > 1) { oz -= g.nc[2]; Sz = -g.box[2]; } else if (oz < -1) { oz += g.nc[2]; Sz = g.box[2]; }
	bool self = (ox == 0 && oy == 0 && oz == 0);
	bool fwd = (oz == 1) || (oz == 0 && (ox == 1 || (ox == 0 && oy == 1)));
	bool home = self ? (i > j) : fwd;
	if (!home && !always) return 0.0;   // always: an unshifted pair the caller visits once (symmetric tables)
	double qx = pj.x, qy = pj.y, qz = pj.z;
	if (!self && !always) { qx += Sx; qy += Sy; qz += Sz; }   // adding a zero shift is exact
	double dx = pi.x - qx, dy = pi.y - qy, dz = pi.z - qz;
	double dr2 = dx * dx + dy * dy + dz * dz;
	double uo = (dr2 < g.rc2) ? pair_potential_val(dr2, pi.type, pj.type, nT, uC) : 0.0;
	if (EMODE == 1) return uo;
	double ex = pi.x * sx - qx * sx, ey = pi.y * sy - qy * sy, ez = pi.z * sz - qz * sz;
	double er2 = ex * ex + ey * ey + ez * ez;
	double un = (er2 < g.rc2) ? pair_potential_val(er2, pi.type, pj.type, nT, uC) : 0.0;
	return uo - un;
}

// 1/sqrt(x) to ~2^-43 from the MUFU.RSQ64H seed and one Newton step; x normal and positive
__device__ __forceinline__ double rsqrt43(double x)
{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	double e = __fma_rn(-(x * y), y, 1.0);
	return __fma_rn(0.5 * y, e, y);
}

// SYMM: both constant tables are symmetric in the two types (checked on the host when they are set; true for every
// generator of the reference).  Then the orientation of a pair only matters through the image shift,
// and an unshifted pair needs nothing but d = p_i - p_j, which is exactly antisymmetric.
//
// Phase-1 cutoff per candidate: min(a[type_i], a[type_j]) with a[] a per-type FP32 radius^2 (+ margin) carried in
// pos32[].w.  a[t] = rc^2 + margin when type t has a partner type with a non-zero tail branch; rm^2 + margin when
// every pair (t,u) has zero tail constants (purely repulsive types, e.g. HEAD: Umin = 0 -- beyond rm their force is
// exactly +-0 in the reference's arithmetic, so the candidates can be dropped unseen); negative when t interacts
// with nothing.  min(a_i, a_j) is never below the exact per-pair cutoff, so nothing that matters is dropped.
//
// ptab (staged to shared): per ordered type pair 10 doubles {T1, T2 | c0, c1, c2, - | c3, c4, c5, -} = the
// reference's row (MD.h:795-848) plus the two exact thresholds on r^2 that decide its branch:
// r < rm <=> r^2 < T1 and r < 2 rm <=> r^2 < T2 (smallest doubles whose correctly rounded square root reaches rm,
// 2 rm; computed on the host), so the branch never waits for the square root.
constexpr int PTAB_STRIDE = 10;

// Work distribution inside a block (128 threads, 128 consecutive slots):
//   * phase 1 is per particle, but lanes of a warp wait for the slowest one and the work differs several-fold between
//     type classes, so particles are dealt to threads class by class (long-range types fill the first warps);
//     every lane then walks ITS OWN flat stream of candidate ranges (a small per-thread table in shared memory),
//     so lanes only meet again at the end of the phase, not after every row.  A list entry is 16 bits:
//     (range index, offset inside the range);
//   * phase 2 is per list.  After a block barrier the lists are handed out again sorted by length (counting sort
//     in shared memory), so the lanes of a warp drain lists of nearly equal length; the thread that drains a list
//     loads that particle's record and writes its acceleration -- no reduction anywhere.
// A list that fills up before phase 1 ends is drained on the spot by its own thread (never seen in practice: 128
// entries); that partial sum travels with the list.  Pairs seen through a periodic image (particles in the outermost
// cell layers only) go straight to the general routine from a separate, plain loop.
constexpr int PAIR_NSEG = 9;

template <int BT>
struct PairSmemT {
	int perm[PAIR_TPB];                 // particle (slot) taken by each thread in phase 1
	int cnt[BT];                        // list length per phase-1 thread
	int order[BT];                      // phase-2 thread -> phase-1 thread whose list it drains
	double part[3][BT];                 // sums that do not go through the list (early drains, periodic images)
	int seg_b[PAIR_NSEG][BT];           // candidate ranges of each thread: first index ...
	unsigned short seg_n[PAIR_NSEG][BT];   // ... and length (< 4096)
	int hist[PAIR_CAP + 2];
	int wcnt[PAIR_TPB / 32];
};
typedef PairSmemT<PAIR_TPB> PairSmem;

//
// EMODE 0: forces.  EMODE 1 / 2: the same two-phase machinery evaluates the pair potential / the dPotential of a box
// move (SYMM tables only; tab = uC, ptab = the potential's padded table): every unordered pair is visited ONCE --
// phase 1 walks only the forward rows of the stencil (oz = 1, or oz = 0 and oy = 1) and, in the particle's own row,
// the slots behind its own (the rest of its cell and the cell at +x) -- and a block reduction replaces the write of
// a[].  Phase 1 widens its cutoffs by `extra32`, the most a component-wise scaling of the box can move r^2 across a
// cutoff (a pair outside rc before the move and inside after it still has a term).
// EMODE 3: forces AND the dPotential of a box move proposed for the same configuration (the Metropolis trial that follows
// an MD step sees the positions of that step's force evaluation, MD.cpp:511-615): the force pass with the widened
// phase-1 cutoffs of EMODE 2; every in-range (before or after the move) entry also contributes half of its U - U', the
// other half comes from the other end of the pair.  Forces stay bit-identical to EMODE 0 (same pairs, same order).
// (EnergyArgs: smd_internal.cuh)

template <int EMODE, bool LANGEVIN, bool SYMM, int SPLIT = 1>
__global__ void __launch_bounds__(PairCfg<SPLIT>::BT, PairCfg<SPLIT>::BLOCKS) k_pair_force2(Cnt cnt, int cap, const Particle *__restrict__ pos,
                                                            const float4 *__restrict__ pos32, const int *__restrict__ start,
                                                            const int *__restrict__ win, Geom g, int nT,
                                                            const double *__restrict__ tab, const double *__restrict__ ptab,
                                                            PairGeo pg, double *__restrict__ acc, LangevinArgs lg,
                                                            const int *__restrict__ gid, EnergyArgs en,
                                                            const uint2 *__restrict__ pos16)
{
	// (its dependents are released further down, see pg.done)  The tail launch of a hybrid pair is a programmatic dependent of
	// the MAIN launch, whose results it does not need: it only needs the cell build, and that had completed and flushed before
	// the first main block got past this very wait -- and no tail block is resident before every main block has started.
	if (!pg.nowait) asm volatile("griddepcontrol.wait;" ::: "memory");
	SMD_TL(3);
	static_assert(EMODE != 3 || SYMM, "forces + dPotential in one pass: symmetric tables only");
	constexpr bool ENERGY_ONLY = (EMODE == 1 || EMODE == 2);   // no forces; every unordered pair once
	constexpr bool DU = (EMODE == 3);                          // forces + dPotential
	// SPLIT = 3: three threads per particle, one per z plane of the stencil (warp w of the block's three takes the rows
	// oz = w - 1 of the same 32 particles): a third of the critical path per thread -- a block of a small system, or of the
	// last, sparse round of blocks of a large one, is done in a third of the time -- and three partial sums per particle,
	// added in plane order by the first warp.  Forces only.
	static_assert(SPLIT == 1 || (SPLIT == 3 && !ENERGY_ONLY), "the split engine evaluates forces");
	constexpr int BT = PairCfg<SPLIT>::BT;        // threads per block
	constexpr int NP = PairCfg<SPLIT>::NP;        // particles per block
	constexpr int CAP = PairCfg<SPLIT>::CAP;      // list entries per thread
	constexpr int NROW = PairCfg<SPLIT>::NROW;    // stencil rows per thread
	static_assert(CAP <= PAIR_CAP, "hist[] is sized for the one-thread engine");
	const int N = cnt.get();
	const int bid = (int)blockIdx.x;
	const int base = pg.first + bid * NP;   // first slot of this block
	if (base >= N) {
		if (EMODE != 0 && threadIdx.x == 0) en.partials[pg.part0 + blockIdx.x] = 0.0;
		return;
	}
	// pg.done: the step seam is launched as a programmatic dependent of this kernel.  All its blocks may become resident as
	// soon as every block of this grid has started (they only fit where pair blocks have left: the tail of the grid); seam
	// block b then waits for done[b] == epoch, which block b of this grid sets when its accelerations are written.
	if (pg.done) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	extern __shared__ __align__(128) unsigned char s_raw[];
	typedef PairSmemT<BT> Smem;
	Smem &sm = *reinterpret_cast<Smem *>(s_raw);
	double *s_ptab = reinterpret_cast<double *>(s_raw + ((sizeof(Smem) + 15) & ~size_t(15)));
	const int nptab = PTAB_STRIDE * nT * nT;
	// EMODE 3 only (its launches reserve the room; every byte of shared memory is L1 the other instances want): the
	// potential's padded table next to the force's, and the dPotential terms that bypass the lists
	double *s_utab = s_ptab + nptab;
	double *s_dup = s_utab + nptab;
	unsigned short *s_lists = reinterpret_cast<unsigned short *>(DU ? s_dup + BT : s_ptab + nptab);
	const int row0 = SPLIT > 1 ? NROW * (int)(threadIdx.x >> 5) : 0;   // first stencil row of this thread (SPLIT = 1)
	// SPLIT = 3: which three rows a warp walks.  By z plane the middle warp got the own row AND two face rows, 64 % of the
	// candidates, and the other two waited for it at the hand-over (barrier stall 2.5 per issue, profiles/r02n_pair3_kernel.md);
	// dealt by weight instead: own row + two corner rows / two face rows + a corner row / the same (r = 3 (oz + 1) + oy + 1).
	const unsigned rowmap = SPLIT == 3 ? ((threadIdx.x >> 5) == 0 ? 0x804u : (threadIdx.x >> 5) == 1 ? 0x231u : 0x675u) : 0u;
	auto row_of = [&](int rr) -> int { return SPLIT == 3 ? (int)((rowmap >> (4 * rr)) & 15u) : row0 + rr; };
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	for (int k = tid; k < nptab; k += BT) s_ptab[k] = ptab[k];
	if (DU) for (int k = tid; k < nptab; k += BT) s_utab[k] = en.utab[k];
	for (int k = tid; k < CAP + 2; k += BT) sm.hist[k] = 0;
	const int w0 = win[WIN_ORG], w1 = win[WIN_ORG + 1], w2 = win[WIN_ORG + 2];
	const int d0 = win[WIN_DIM], d1 = win[WIN_DIM + 1], d2 = win[WIN_DIM + 2];
	const int fd0 = win[WIN_FD0], xs = g.xs;


	SMD_PC_INIT;
	// ---- deal the block's particles to threads by class
	{
		const int pt = SPLIT > 1 ? lane : tid;   // (split engines: every warp holds the same 32 particles)
		const int i0 = base + pt;
		const bool heavy = (i0 < N) && pos32[i0].w >= pg.thr32;
		const unsigned bal = __ballot_sync(0xffffffffu, heavy);
		if (lane == 0) sm.wcnt[wid] = __popc(bal);
		__syncthreads();
		int before = 0, total = 0;
#pragma unroll
		for (int k = 0; k < NP / 32; k++) { int c = sm.wcnt[k]; total += c; if (k < (SPLIT > 1 ? 0 : wid)) before += c; }
		const int below = __popc(bal & ((1u << lane) - 1u));
		const int rank = heavy ? before + below : total + (pt - before - below);
		if (SPLIT == 1 || wid == 0) sm.perm[rank] = i0;
		__syncthreads();
	}
	const int i = sm.perm[SPLIT > 1 ? lane : tid];
	// slab mode: ghosts are only neighbours, nobody gathers for them
	const bool live = i < N && !(g.slab && (gid[i] & GID_GHOST));
	Particle pi;
	pi.x = pi.y = pi.z = 0; pi.type = 0; pi.cell = 0;
	float4 p32 = make_float4(0.f, 0.f, 0.f, -1.f);
	int cx = 0, cy = 0, cz = 0;
	if (live) {
		pi = load_particle(pos + i);
		p32 = pos32[i];
		unpack_cell(pi.cell, cx, cy, cz);
	}
	const double rc2 = g.rc2;
	const float ai = p32.w;
	// entries of thread t's list sit 64 B apart (one 16-bit column per lane)
	auto list_base = [&](int t) { return (unsigned)__cvta_generic_to_shared(s_lists + (t >> 5) * (CAP * 32) + (t & 31)); };
	const unsigned segb_base = (unsigned)__cvta_generic_to_shared(&sm.seg_b[0][0]);

	double du = 0.0;   // EMODE 3: this thread's share of the dPotential
	// phase 2 for one list: the particle (slot io, record po) whose ranges belong to phase-1 thread t, against the
	// entries [rp, wend).  Branch-free except for the hand-over to the general routine (asymmetric tables, r >= 2 rm),
	// so that two pairs interleave; the records of the next two pairs are already in flight.  sqrt and the division
	// share one reciprocal square root, each finished with an exact-residual correction step (correctly rounded except
	// for vanishingly rare near-ties, then off by one ulp).
	auto drain = [&](int io, const Particle &po, int t, unsigned rp, unsigned wend, double &ax, double &ay, double &az) {
		const char *rowi = reinterpret_cast<const char *>(s_ptab + PTAB_STRIDE * po.type * nT);
		const unsigned tb = segb_base + 4u * (unsigned)t;
		auto fetch = [&](unsigned a, int &j) {       // entry -> neighbour slot and record
			unsigned short e;
			asm volatile("ld.shared.u16 %0, [%1];" : "=h"(e) : "r"(a) : "memory");
			int b;
			asm volatile("ld.shared.s32 %0, [%1];" : "=r"(b) : "r"(tb + (unsigned)(e >> PAIR_SEGBITS) * (4u * BT)) : "memory");
			j = b + (int)(e & ((1u << PAIR_SEGBITS) - 1u));
			return load_particle(pos + j);
		};
		// Potential<T>, MD.h:895-930, on r^2 (x normal and positive): correctly rounded sqrt as above, both branches
		// evaluated and selected
		auto upoly = [&](double dr, const char *c) {   // the potential at distance dr (= the correctly rounded sqrt of r^2)
			double2 c01 = *reinterpret_cast<const double2 *>(c + 16);
			double c2 = *reinterpret_cast<const double *>(c + 32);
			double2 c34 = *reinterpret_cast<const double2 *>(c + 48);
			double c5 = *reinterpret_cast<const double *>(c + 64);
			double tc = c01.x - dr, tt = c34.x - dr;
			double ucore = c01.y * tc * tc + c2;
			double utail = tt * tt * (c34.y - tt * c5);
			return (dr <= c01.x) ? ucore : utail;
		};
		auto upot = [&](double x, const char *c) {
			double y = rsqrt43(x);
			double dr = x * y;
			dr = __fma_rn(__fma_rn(-dr, dr, x), 0.5 * y, dr);
			double2 c01 = *reinterpret_cast<const double2 *>(c + 16);
			double c2 = *reinterpret_cast<const double *>(c + 32);
			double2 c34 = *reinterpret_cast<const double2 *>(c + 48);
			double c5 = *reinterpret_cast<const double *>(c + 64);
			double tc = c01.x - dr, tt = c34.x - dr;
			double ucore = c01.y * tc * tc + c2;
			double utail = tt * tt * (c34.y - tt * c5);
			return (dr <= c01.x) ? ucore : utail;
		};
		const char *rowu = reinterpret_cast<const char *>(s_utab + PTAB_STRIDE * po.type * nT);   // EMODE 3
		auto fast = [&](int j, const Particle &pj) -> bool {
			double dx = po.x - pj.x, dy = po.y - pj.y, dz = po.z - pj.z;
			double dr2 = dx * dx + dy * dy + dz * dz;
			const char *c = rowi + (PTAB_STRIDE * 8) * pj.type;
			double2 T = *reinterpret_cast<const double2 *>(c);
			bool in = dr2 < rc2 && j != io;            // the particle itself passes phase 1 (r2 = 0)
			bool ok = SYMM && in && dr2 < T.y;
			c += (dr2 < T.x) ? 16 : 48;                  // pair_branch(): core or tail constants
			double2 c01 = *reinterpret_cast<const double2 *>(c);
			double c2 = *reinterpret_cast<const double *>(c + 16);
			double x = ok ? dr2 : 1.0;
			double y = rsqrt43(x);
			double dr = x * y;
			dr = __fma_rn(__fma_rn(-dr, dr, x), 0.5 * y, dr);     // sqrt(x)
			double m = c01.x - dr;
			double num = (c01.y - c2 * m) * m;
			double q = num * y;
			q = __fma_rn(__fma_rn(-dr, q, num), y, q);            // num / dr
			q = ok ? q : 0.0;
			ax += dx * q; ay += dy * q; az += dz * q;
			if (DU && !(in && !ok)) {   // (pairs handed to the general routine take their energy term there)
				const char *cu = rowu + (PTAB_STRIDE * 8) * pj.type;
				// (here the pair is either out of range or took the fast path: dr above IS sqrt(dr2), the very value upot would
				// compute again -- same operations on the same operand)
				double uo = upoly(dr, cu);
				uo = in ? uo : 0.0;
				double ex = po.x * en.sx - pj.x * en.sx, ey = po.y * en.sy - pj.y * en.sy, ez = po.z * en.sz - pj.z * en.sz;
				double er2 = ex * ex + ey * ey + ez * ez;
				bool in2 = er2 < rc2 && j != io;
				double un = upot(in2 ? er2 : 1.0, cu);
				du += 0.5 * (uo - (in2 ? un : 0.0));   // the other half: the same entry in the neighbour's list
			}
			return in && !ok;
		};
		auto general = [&](int j, const Particle &pj) {
			D3 f = pair_force_term(io, po, j, pj, g, nT, tab, 6 * nT * nT);
			ax += f.x; ay += f.y; az += f.z;
			if (DU) du += pair_energy_term<2>(io, po, j, pj, g, nT, en.uC, en.sx, en.sy, en.sz, false);   // counted by the pair's home particle
		};
		auto efast = [&](int j, const Particle &pj) {
			double dx = po.x - pj.x, dy = po.y - pj.y, dz = po.z - pj.z;
			double dr2 = dx * dx + dy * dy + dz * dz;
			const char *c = rowi + (PTAB_STRIDE * 8) * pj.type;
			bool in = dr2 < rc2 && j != io;
			double u = upot(in ? dr2 : 1.0, c);
			u = in ? u : 0.0;
			if (EMODE == 2) {
				double ex = po.x * en.sx - pj.x * en.sx, ey = po.y * en.sy - pj.y * en.sy, ez = po.z * en.sz - pj.z * en.sz;
				double er2 = ex * ex + ey * ey + ez * ez;
				bool in2 = er2 < rc2 && j != io;
				double un = upot(in2 ? er2 : 1.0, c);
				u = u - (in2 ? un : 0.0);
			}
			ax += u;
		};
		if (ENERGY_ONLY) {
			while (rp < wend) {
				int j0;
				Particle p0 = fetch(rp, j0);
				rp += 64u;
				if (rp < wend) {
					int j1;
					Particle p1 = fetch(rp, j1);
					rp += 64u;
					efast(j0, p0);
					efast(j1, p1);
				} else {
					efast(j0, p0);
				}
			}
			return;
		}
		if (rp + 64u < wend) {
			int j0, j1;
			Particle p0 = fetch(rp, j0), p1 = fetch(rp + 64u, j1);
			rp += 128u;
			while (true) {
				int n0 = j0, n1 = j1;
				Particle q0 = p0, q1 = p1;
				const bool more = rp + 64u < wend;
				if (more) { q0 = fetch(rp, n0); q1 = fetch(rp + 64u, n1); }   // next two pairs: in flight during the math below
				bool s0 = fast(j0, p0);
				bool s1 = fast(j1, p1);
				if (s0 || s1) {
					if (s0) general(j0, p0);
					if (s1) general(j1, p1);
				}
				if (!more) break;
				rp += 128u;
				j0 = n0; j1 = n1; p0 = q0; p1 = q1;
			}
		}
		if (rp < wend) {
			int j0;
			Particle p0 = fetch(rp, j0);
			if (fast(j0, p0)) general(j0, p0);
		}
	};

	const unsigned lbase = list_base(tid);
	unsigned wp = lbase;                         // shared-window address of the next free entry of this thread's list
	double ex = 0, ey = 0, ez = 0;               // sums that bypass the list
	auto push = [&](unsigned v) {
		asm volatile("st.shared.u16 [%0], %1;" ::"r"(wp), "h"((unsigned short)v) : "memory");
		wp += 64u;
	};
	unsigned wlim = lbase + 64u * (CAP - 8);    // checked once per two groups of four
	asm volatile("" : "+r"(wlim));               // opaque: rematerialising the shared-window address cost 8 instructions per group

	SMD_PC(0);   // entry: table staging, dealing, own record
	// ---- the particle's candidate ranges: one per (y,z) row of the stencil, pruned by geometry
	// conservative FP32 distances to the faces of the own cell: a neighbour cell at offset -1 / +1 along an axis holds
	// no point closer than that along the axis
	float fm[3], fp[3];
	{
		float c[3] = {p32.x, p32.y, p32.z};
		int ci[3] = {cx, cy, cz};
#pragma unroll
		for (int a = 0; a < 3; a++) {
			fm[a] = fmaxf(c[a] - (float)ci[a] * pg.cs32[a] - pg.slack32, 0.f);
			fp[a] = fmaxf((float)(ci[a] + 1) * pg.cs32[a] - c[a] - pg.slack32, 0.f);
		}
	}
	const float ext = EMODE != 0 ? en.extra32 : 0.f;
	const float amax = fminf(ai, pg.thr32) + ext;
	int nseg = 0;
	bool shifted_rows = false;
	int rjb[NROW], rje[NROW];
	// all the look-ups of start[] are issued before any of them is used (one trip to L2 instead of nine)
#pragma unroll
	for (int rr = 0; rr < NROW; rr++) {
		const int r = row_of(rr);
		const int oz = r / 3 - 1, oy = r - 3 * (r / 3) - 1;
		int nz = cz + oz, ny = cy + oy;
		bool wrapyz = nz < 0 || nz >= g.nc[2] || ny < 0 || ny >= g.nc[1];
		int lz = nz - w2, ly = ny - w1;
		float gy = oy == 0 ? 0.f : (oy > 0 ? fp[1] : fm[1]), gz = oz == 0 ? 0.f : (oz > 0 ? fp[2] : fm[2]);
		float gyz = gy * gy + gz * gz;
		bool row_ok = live && gyz < amax;
		if (row_ok && (wrapyz || cx == 0 || cx == g.nc[0] - 1)) shifted_rows = true;
		if (ENERGY_ONLY && (oz < 0 || (oz == 0 && oy < 0))) row_ok = false;   // backward rows: the other particle counts the pair
		row_ok = row_ok && !wrapyz && lz >= 0 && lz < d2 && ly >= 0 && ly < d1;
		bool keep_lo = gyz + fm[0] * fm[0] < amax, keep_hi = gyz + fp[0] * fp[0] < amax;
		// cells cx-1 .. cx+1 that need no wrap, clamped to the window (cells outside it are empty)
		int xlo, xhi;
		if (g.slab) {   // the window holds every neighbour column of an owned particle, possibly across the periodic seam
			xlo = win_x(max(cx - (keep_lo ? 1 : 0), 0), w0, g.nc[0]); xhi = win_x(min(cx + (keep_hi ? 1 : 0), g.nc[0] - 1), w0, g.nc[0]);
		} else {
			xlo = max(max(cx - (keep_lo ? 1 : 0), 0) - w0, 0); xhi = min(min(cx + (keep_hi ? 1 : 0), g.nc[0] - 1) - w0, d0 - 1);
		}
		rjb[rr] = 0; rje[rr] = 0;
		// x slices: the row is sorted by x to cs / xs, read only the slices within reach of this particle in this row
		// (conservative: the FP32 errors of the slice coordinate are ~1e-4 of a slice, the slack is 0.02)
		int flo = xlo * xs, fhi = xhi * xs + (xs - 1);
		if (xs > 1) {
			const float reach = sqrtf(fmaxf(amax - gyz, 0.f)) + pg.slack32;
			// x relative to the window origin: own window column (wraps across the periodic seam in slab mode) + offset in the cell
			const float rel = (float)win_x(cx, w0, g.nc[0]) * pg.cs32[0] + (p32.x - (float)cx * pg.cs32[0]);
			flo = max(flo, (int)floorf((rel - reach) * pg.finv32 - 0.02f));
			fhi = min(fhi, (int)floorf((rel + reach) * pg.finv32 + 0.02f));
		}
		if (row_ok && flo <= fhi) {
			int rowbase = fd0 * (ly + d1 * lz);
			rjb[rr] = start[rowbase + flo]; rje[rr] = start[rowbase + fhi + 1];
		}
	}
#pragma unroll
	for (int rr = 0; rr < NROW; rr++) {
		const int r = row_of(rr);
		int jb = rjb[rr];
		const int je = rje[rr];
		if (ENERGY_ONLY && r == 4) jb = min(max(jb, i + 1), je);   // own row: only the slots behind mine
		const int lim = (1 << PAIR_SEGBITS) - 4;
		sm.seg_b[r][tid] = jb;
		sm.seg_n[r][tid] = (unsigned short)min(je - jb, lim);
		// a range longer than the 12-bit offset field (> 1300 particles per cell): take the excess one by one
		for (int j = jb + lim; j < je; j++) {
			float4 c = pos32[j];
			float dx = p32.x - c.x, dy = p32.y - c.y, dz = p32.z - c.z;
			if (__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)) < fminf(ai, c.w) + ext && j != i) {
				if (!ENERGY_ONLY) {
					const Particle pj = load_particle(pos + j);
					D3 f = pair_force_term(i, pi, j, pj, g, nT, tab, 6 * nT * nT);
					ex += f.x; ey += f.y; ez += f.z;
					if (DU) du += pair_energy_term<2>(i, pi, j, pj, g, nT, en.uC, en.sx, en.sy, en.sz, false);
				} else {
					ex += pair_energy_term<(ENERGY_ONLY ? EMODE : 1)>(i, pi, j, load_particle(pos + j), g, nT, tab, en.sx, en.sy, en.sz, true);
				}
			}
		}
		nseg = (je > jb) ? rr + 1 : nseg;
	}

	SMD_PC(1);   // range set-up
	// ---- phase 1: prefilter along the thread's own stream of ranges over the 8-byte candidate records (16-bit
	// window-relative coordinates, see quantize16) -- from the block's staged copy in shared memory where the range was
	// staged, else from global memory; four candidates per step, the next four already in flight.  The coordinates are
	// spliced into the mantissa of 2^23 (one PRMT each), so that a float subtraction gives their exact difference.
	float qx = 0.f, qy = 0.f, qz = 0.f, aiq = -INFINITY, extq = 0.f;
	if (live) {
		const uint2 me = pos16[i];
		qx = __uint_as_float(__byte_perm(me.x, 0x4B000000u, 0x7610));
		qy = __uint_as_float(__byte_perm(me.x, 0x4B000000u, 0x7632));
		qz = __uint_as_float(__byte_perm(me.y, 0x4B000000u, 0x7610));
		aiq = __uint_as_float(me.y & 0xffff0000u);
	}
	if (EMODE != 0) {
		// the widening of the cutoffs in steps^2: (R^2 + ext) / res^2 plus the cross term of the 1.75-step slack
		const float ir = __int_as_float(win[WIN_INVRES]);
		extq = (ext * ir) * ir * (1.000001f + 1.75f / (pg.rmin32 * ir));
	}
	const int aiqi = __float_as_int(aiq);
	for (int sq = 0; sq < nseg; sq++) {
		const int sg = row_of(sq);
		const int n = sm.seg_n[sg][tid];
		if (n == 0) continue;
		const uint2 *cp = pos16 + sm.seg_b[sg][tid];
		const unsigned tag = (unsigned)sg << PAIR_SEGBITS;
		uint2 ga[4], gb[4];                          // ping-pong buffers: one group under test, the next in flight
#pragma unroll
		for (int k = 0; k < 4; k++) ga[k] = cp[k];   // both arrays are padded: the overhang is masked below
		int q = 0;
		// tests the group c[] at offset q and meanwhile loads the following one into nx[]; false after the last group
		auto group = [&](const uint2 (&c)[4], uint2 (&nx)[4]) {
			const unsigned e0 = tag | (unsigned)q;
			const int rem = n - q;
			q += 4;
			const bool more = q < n;
			if (more) {
#pragma unroll
				for (int k = 0; k < 4; k++) nx[k] = cp[q + k];
			}
#pragma unroll
			for (int k = 0; k < 4; k++) {
				float dx = qx - __uint_as_float(__byte_perm(c[k].x, 0x4B000000u, 0x7610));
				float dy = qy - __uint_as_float(__byte_perm(c[k].x, 0x4B000000u, 0x7632));
				float dz = qz - __uint_as_float(__byte_perm(c[k].y, 0x4B000000u, 0x7610));
				float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
				if (EMODE != 0) {
					const float cw = __uint_as_float(c[k].y & 0xffff0000u);
					if (r2 < fminf(aiq, cw) + extq && k < rem) push(e0 + k);
				} else {
					// r2 >= 0, so its bit pattern orders like an integer; the candidate's cutoff is compared in place, with
					// the z half-word below it (less than one bf16 step more permissive; phase 2 sorts it out)
					const int r2i = __float_as_int(r2);
					if (r2i < (int)c[k].y && r2i < aiqi && k < rem) push(e0 + k);
				}
			}
			return more;
		};
		while (true) {
			if (!group(ga, gb)) break;
			if (!group(gb, ga)) break;
			if (wp > wlim) { drain(i, pi, tid, lbase, wp, ex, ey, ez); wp = lbase; }   // list nearly full: never seen in practice
		}
		if (wp > wlim) { drain(i, pi, tid, lbase, wp, ex, ey, ez); wp = lbase; }
	}


	SMD_PC(2);   // phase 1
	// ---- rows and end cells seen through a periodic image (particles in the outermost cell layers only)
	if (shifted_rows) {
#pragma unroll 1
		for (int s = 0; s < 3 * NROW; s++) {
			int r = row_of(s / 3), sub = s - 3 * (s / 3);
			int oz = r / 3 - 1, oy = r - 3 * (r / 3) - 1;
			int nz = cz + oz, ny = cy + oy;
			float sx = 0.f, sy = 0.f, sz = 0.f;
			if (nz < 0) { nz += g.nc[2]; sz = -(float)g.box[2]; }
			if (nz >= g.nc[2]) { nz -= g.nc[2]; sz = (float)g.box[2]; }
			if (ny < 0) { ny += g.nc[1]; sy = -(float)g.box[1]; }
			if (ny >= g.nc[1]) { ny -= g.nc[1]; sy = (float)g.box[1]; }
			int lz = nz - w2, ly = ny - w1;
			float gy = oy == 0 ? 0.f : (oy > 0 ? fp[1] : fm[1]), gz = oz == 0 ? 0.f : (oz > 0 ? fp[2] : fm[2]);
			float gyz = gy * gy + gz * gz;
			bool row_ok = gyz < amax && lz >= 0 && lz < d2 && ly >= 0 && ly < d1;
			bool keep_lo = gyz + fm[0] * fm[0] < amax, keep_hi = gyz + fp[0] * fp[0] < amax;
			int xlo, xhi;
			if (sub == 0) {          // the unwrapped x range of a row shifted in y or z (unshifted rows were done above)
				if (g.slab) {
					xlo = win_x(max(cx - (keep_lo ? 1 : 0), 0), w0, g.nc[0]); xhi = win_x(min(cx + (keep_hi ? 1 : 0), g.nc[0] - 1), w0, g.nc[0]);
				} else {
					xlo = max(max(cx - (keep_lo ? 1 : 0), 0) - w0, 0); xhi = min(min(cx + (keep_hi ? 1 : 0), g.nc[0] - 1) - w0, d0 - 1);
				}
				row_ok = row_ok && (sy != 0.f || sz != 0.f);
			} else if (sub == 1) {   // left face: the image of the last cell of the row
				xlo = xhi = win_x(g.nc[0] - 1, w0, g.nc[0]); sx = -(float)g.box[0];
				row_ok = row_ok && keep_lo && cx == 0 && xlo < d0;
			} else {                 // right face: the image of the first cell
				xlo = xhi = win_x(0, w0, g.nc[0]); sx = (float)g.box[0];
				row_ok = row_ok && keep_hi && cx == g.nc[0] - 1 && xlo < d0;
			}
			if (!row_ok || xlo > xhi) continue;
			int rowbase = fd0 * (ly + d1 * lz);
			int jb = start[rowbase + xlo * xs], je = start[rowbase + (xhi + 1) * xs];
			const float qx = p32.x - sx, qy = p32.y - sy, qz = p32.z - sz;
			for (int j = jb; j < je; j++) {
				float4 c = pos32[j];
				float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
				if (__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)) < fminf(ai, c.w) + ext) {
					if (!ENERGY_ONLY) {
						const Particle pj = load_particle(pos + j);
						D3 f = pair_force_term(i, pi, j, pj, g, nT, tab, 6 * nT * nT);
						ex += f.x; ey += f.y; ez += f.z;
						if (DU) du += pair_energy_term<2>(i, pi, j, pj, g, nT, en.uC, en.sx, en.sy, en.sz, false);
					} else {
						ex += pair_energy_term<(ENERGY_ONLY ? EMODE : 1)>(i, pi, j, load_particle(pos + j), g, nT, tab, en.sx, en.sy, en.sz, false);
					}
				}
			}
		}
	}

	SMD_PC(3);   // periodic-image rows
	// ---- hand the lists out again, longest first
	const int lcnt = (int)((wp - lbase) >> 6);
	sm.cnt[tid] = lcnt;
	sm.part[0][tid] = ex; sm.part[1][tid] = ey; sm.part[2][tid] = ez;
	if (DU) s_dup[tid] = du;   // energy terms that bypassed the list (periodic images, early drains): they travel with it
	atomicAdd(&sm.hist[lcnt], 1);
	__syncthreads();
	if (tid < 32) {   // exclusive prefix over descending length (CAP + 1 bins)
		constexpr int PER = (CAP + 1 + 31) / 32;
		int h[PER], sum = 0;
#pragma unroll
		for (int k = 0; k < PER; k++) { int c = CAP - (tid * PER + k); h[k] = c >= 0 ? sm.hist[c] : 0; sum += h[k]; }
		int inc = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, d); if (tid >= d) inc += v; }
		int run = inc - sum;
#pragma unroll
		for (int k = 0; k < PER; k++) { int c = CAP - (tid * PER + k); if (c >= 0) sm.hist[c] = run; run += h[k]; }
	}
	__syncthreads();
	sm.order[atomicAdd(&sm.hist[lcnt], 1)] = tid;
	__syncthreads();

	SMD_PC(4);   // hand-over: histogram, prefix, order (three barriers: includes waiting for the block's slowest warp)
	// ---- phase 2: drain one list, FP64
	const int o = sm.order[tid];
	const int io = sm.perm[SPLIT > 1 ? (o & 31) : o];
	const bool act = io < N && !(g.slab && (gid[io] & GID_GHOST));
	if (SPLIT == 1 && EMODE == 0 && !act && !pg.done) return;
	double ax = sm.part[0][o], ay = sm.part[1][o], az = sm.part[2][o];
	if (DU) du = s_dup[o];
	if (act) {
		const Particle po = load_particle(pos + io);
		const unsigned ob = list_base(o);
		drain(io, po, o, ob, ob + 64u * (unsigned)sm.cnt[o], ax, ay, az);
	}
	if (DU) {   // the block's share of the dPotential, summed in particle order (see below); then the force epilogue
		__syncthreads();
		sm.part[0][o] = act ? du : 0.0;
		__syncthreads();
		double tot = block_sum(sm.part[0][tid]);
		if (tid == 0) en.partials[pg.part0 + blockIdx.x] = tot;
		if (SPLIT == 1 && !act && !pg.done) return;
	}
	if (ENERGY_ONLY) {   // one partial sum per block, reduced deterministically by k_final_sum
		// summed in particle order, not in the (arrival-dependent) order the lists were handed out in: the energy is
		// reproducible to the last bit
		__syncthreads();
		sm.part[0][o] = act ? ax : 0.0;
		__syncthreads();
		double tot = block_sum(sm.part[0][tid]);
		if (tid == 0) en.partials[pg.part0 + blockIdx.x] = tot;
		return;
	}
	int iw = io;        // the particle whose acceleration this thread writes
	bool actw = act;
	if (SPLIT > 1) {   // the partial sums of a particle, added in row order by the first warp
		__syncthreads();
		sm.part[0][o] = act ? ax : 0.0; sm.part[1][o] = act ? ay : 0.0; sm.part[2][o] = act ? az : 0.0;
		__syncthreads();
		iw = sm.perm[lane];
		actw = wid == 0 && iw < N && !(g.slab && (gid[iw] & GID_GHOST));
		if (!actw && !pg.done) return;
		if (actw) {
			ax = sm.part[0][lane]; ay = sm.part[1][lane]; az = sm.part[2][lane];
#pragma unroll
			for (int k = 1; k < SPLIT; k++) { ax += sm.part[0][32 * k + lane]; ay += sm.part[1][32 * k + lane]; az += sm.part[2][32 * k + lane]; }
		}
	}
	if (!actw) {
		// (only reached with pg.done: every thread of the block takes part in the hand-over below)
	} else if (LANGEVIN) {
		int id = lg.gid[iw] & GID_MASK;
		double u[3];
		if (lg.ext_noise) {
			u[0] = lg.ext_noise[3 * id]; u[1] = lg.ext_noise[3 * id + 1]; u[2] = lg.ext_noise[3 * id + 2];
		} else {
			philox_uniform3(lg.seed, lg.step, (uint32_t)id, u);
		}
		double lx = -lg.gamma * lg.vel[iw] + lg.sigma * (2.0 * u[0] - 1.0);
		double ly = -lg.gamma * lg.vel[cap + iw] + lg.sigma * (2.0 * u[1] - 1.0);
		double lz = -lg.gamma * lg.vel[2 * cap + iw] + lg.sigma * (2.0 * u[2] - 1.0);
		acc[iw] = lx + ax; acc[cap + iw] = ly + ay; acc[2 * cap + iw] = lz + az;
	} else {
		acc[iw] += ax; acc[cap + iw] += ay; acc[2 * cap + iw] += az;
	}
	SMD_PC(5);   // phase 2 + epilogue (warps that returned early are not counted)
	if (pg.done) {   // this block's accelerations are complete: release its seam block
		__threadfence();
		__syncthreads();
		// (one word per 32 slots, whatever the engine: the seam waits for the four words of its 128 slots)
		if (tid < NP / 32) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(pg.done + base / 32 + tid), "r"(pg.epoch) : "memory");
	}
}

template <int MODE>
__global__ void __launch_bounds__(TPB) k_chain(int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of, Geom g,
                                               ChainBlock cb, double *acc, double *partials, double sx, double sy, double sz)
{
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	double usum = 0;
	if (k < cb.nChains) {
		int base = cb.start + k * cb.len;
		int s1 = slot_of[base], s2 = slot_of[base + 1];
		Particle p1 = load_particle(pos + s1), p2 = load_particle(pos + s2);
		V3 a1 = {0, 0, 0}, a2 = {0, 0, 0};
		for (int l = 0; l <= cb.len - 3; l++) {
			bool tail = (l == cb.len - 3);
			int s3 = slot_of[base + l + 2];
			Particle p3 = load_particle(pos + s3);
			V3 da = diff_mi(p1, p2, g), db = diff_mi(p2, p3, g);
			V3 a3 = {0, 0, 0};
			if (MODE == 0) {
				V3 f = harmonic_f(da, cb.c[0], cb.c[1]);
				a1.x += f.x; a1.y += f.y; a1.z += f.z;
				a2.x -= f.x; a2.y -= f.y; a2.z -= f.z;
				if (tail) {
					V3 f2 = harmonic_f(db, cb.c[0], cb.c[1]);
					a2.x += f2.x; a2.y += f2.y; a2.z += f2.z;
					a3.x -= f2.x; a3.y -= f2.y; a3.z -= f2.z;
				}
				V3 fa, fb;
				bend_f(da, db, cb.c[2], cb.c[3], fa, fb);
				a1.x += fa.x; a1.y += fa.y; a1.z += fa.z;
				a2.x += (fb.x - fa.x); a2.y += (fb.y - fa.y); a2.z += (fb.z - fa.z);
				a3.x -= fb.x; a3.y -= fb.y; a3.z -= fb.z;
				// particle 1 of this triplet is complete
				acc[s1] += a1.x; acc[cap + s1] += a1.y; acc[2 * cap + s1] += a1.z;
				if (tail) {
					acc[s2] += a2.x; acc[cap + s2] += a2.y; acc[2 * cap + s2] += a2.z;
					acc[s3] += a3.x; acc[cap + s3] += a3.y; acc[2 * cap + s3] += a3.z;
				}
			} else if (MODE == 1) {
				usum += harmonic_p(da, cb.c[0], cb.c[1]);
				if (tail) usum += harmonic_p(db, cb.c[0], cb.c[1]);
				usum += bend_p(da, db, cb.c[2], cb.c[3]);
			} else {
				double uo = harmonic_p(da, cb.c[0], cb.c[1]);
				if (tail) uo += harmonic_p(db, cb.c[0], cb.c[1]);
				uo += bend_p(da, db, cb.c[2], cb.c[3]);
				V3 ea = scaled(da, sx, sy, sz), eb = scaled(db, sx, sy, sz);
				double un = harmonic_p(ea, cb.c[0], cb.c[1]);
				if (tail) un += harmonic_p(eb, cb.c[0], cb.c[1]);
				un += bend_p(ea, eb, cb.c[2], cb.c[3]);
				usum += (uo - un);
			}
			s1 = s2; s2 = s3; p1 = p2; p2 = p3; a1 = a2; a2 = a3;
		}
	}
	if (MODE != 0) {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
}

// ---- slab mode (multi-GPU) variant of k_chain.  A chain can straddle the slab seam, and which of its members are
// local (owned, or ghost copies inside the halo) changes every step, so the work is found from the particles: one
// thread per local slot; the thread of the lowest-numbered OWNED member of a chain ("leader") walks the chain in the
// reference's triplet order.  A triplet is evaluated when one of its three members is owned here (all three must
// then be local: the halo is two cell columns wide, a triplet spans at most two bonds) and only owned members are
// written, so every owned particle receives exactly the terms, in exactly the order, of the single-GPU kernel.
// Energies (MODE 1, 2): a triplet's terms are counted by the rank that owns its FIRST member -- once globally.
template <int MODE>
__global__ void __launch_bounds__(TPB) k_chain_slab(Cnt cnt, int cap, const Particle *__restrict__ pos, const int *__restrict__ gid,
                                                    const int *__restrict__ slot_of, Geom g, ChainBlock cb, double *acc, double *partials,
                                                    double sx, double sy, double sz, int *errflag)
{
	const int N = cnt.get();
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	double usum = 0;
	bool work = false;
	int base = 0, lead = 0;
	if (s < N) {
		int gi = gid[s];
		int rel = gi - cb.start;   // a ghost's flag bit makes rel huge: excluded below
		if (!(gi & GID_GHOST) && rel >= 0 && rel < cb.nChains * cb.len) {
			int k = rel / cb.len;
			lead = rel - k * cb.len;
			base = cb.start + k * cb.len;
			work = true;
			for (int m = 0; m < lead; m++) {
				bool ow;
				if (slab_find(slot_of, gid, N, base + m, ow) >= 0 && ow) { work = false; break; }
			}
		}
	}
	if (work) {
		// triplets that can contain an owned member start at max(lead - 2, 0)
		int l0 = max(lead - 2, 0);
		bool o1, o2, o3;
		int s1 = slab_find(slot_of, gid, N, base + l0, o1), s2 = slab_find(slot_of, gid, N, base + l0 + 1, o2);
		Particle p1, p2, p3;
		if (s1 >= 0) p1 = load_particle(pos + s1);
		if (s2 >= 0) p2 = load_particle(pos + s2);
		V3 a1 = {0, 0, 0}, a2 = {0, 0, 0};
		for (int l = l0; l <= cb.len - 3; l++) {
			bool tail = (l == cb.len - 3);
			int s3 = slab_find(slot_of, gid, N, base + l + 2, o3);
			if (s3 >= 0) p3 = load_particle(pos + s3);
			V3 a3 = {0, 0, 0};
			bool need = (MODE == 0) ? (o1 || o2 || o3) : o1;
			if (need && (s1 < 0 || s2 < 0 || s3 < 0)) { atomicOr(errflag, ERR_SLAB_MISSING); need = false; }
			if (need) {
				V3 da = diff_mi(p1, p2, g), db = diff_mi(p2, p3, g);
				if (MODE == 0) {
					V3 f = harmonic_f(da, cb.c[0], cb.c[1]);
					a1.x += f.x; a1.y += f.y; a1.z += f.z;
					a2.x -= f.x; a2.y -= f.y; a2.z -= f.z;
					if (tail) {
						V3 f2 = harmonic_f(db, cb.c[0], cb.c[1]);
						a2.x += f2.x; a2.y += f2.y; a2.z += f2.z;
						a3.x -= f2.x; a3.y -= f2.y; a3.z -= f2.z;
					}
					V3 fa, fb;
					bend_f(da, db, cb.c[2], cb.c[3], fa, fb);
					a1.x += fa.x; a1.y += fa.y; a1.z += fa.z;
					a2.x += (fb.x - fa.x); a2.y += (fb.y - fa.y); a2.z += (fb.z - fa.z);
					a3.x -= fb.x; a3.y -= fb.y; a3.z -= fb.z;
				} else if (MODE == 1) {
					usum += harmonic_p(da, cb.c[0], cb.c[1]);
					if (tail) usum += harmonic_p(db, cb.c[0], cb.c[1]);
					usum += bend_p(da, db, cb.c[2], cb.c[3]);
				} else {
					double uo = harmonic_p(da, cb.c[0], cb.c[1]);
					if (tail) uo += harmonic_p(db, cb.c[0], cb.c[1]);
					uo += bend_p(da, db, cb.c[2], cb.c[3]);
					V3 ea = scaled(da, sx, sy, sz), eb = scaled(db, sx, sy, sz);
					double un = harmonic_p(ea, cb.c[0], cb.c[1]);
					if (tail) un += harmonic_p(eb, cb.c[0], cb.c[1]);
					un += bend_p(ea, eb, cb.c[2], cb.c[3]);
					usum += (uo - un);
				}
			}
			if (MODE == 0) {
				if (o1) { acc[s1] += a1.x; acc[cap + s1] += a1.y; acc[2 * cap + s1] += a1.z; }
				if (tail) {
					if (o2) { acc[s2] += a2.x; acc[cap + s2] += a2.y; acc[2 * cap + s2] += a2.z; }
					if (o3) { acc[s3] += a3.x; acc[cap + s3] += a3.y; acc[2 * cap + s3] += a3.z; }
				}
			}
			s1 = s2; s2 = s3; p1 = p2; p2 = p3; o1 = o2; o2 = o3; a1 = a2; a2 = a3;
		}
	}
	if (MODE != 0) {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
}

// ------------------------------------------------------------------------------------------------ slab exchange
// After the particles of a slab moved (Verlet::first, or an accepted box move) and were re-tagged with their cells:
//   k_slab_pack    every rank, one pass over its slots: old ghosts are marked dead; owned particles whose new column
//                  lies outside [col_lo, col_hi) MIGRATE to the neighbour on that side (record + velocity [+ unwrapped
//                  position]) and stay behind as ghosts (their column is inside this rank's halo); owned particles in
//                  the outermost `halo` columns are copied to the neighbour as ghosts.  Entries are written straight
//                  into the neighbour's receive buffer through peer memory; the last block to finish publishes
//                  {count, seq} in the header (release at system scope).
//   k_slab_unpack  waits for both headers of this exchange (acquire), appends the received particles behind the
//                  local ones and publishes the extended count; the build that follows sorts them into place and
//                  drops the dead records.
// One message per neighbour and step: the migrants a rank sends are exactly the ghosts it would otherwise have to
// ask back, and what arrives from the neighbour completes the halo.
__device__ __forceinline__ void st_entry(SlabMsgEntry *e, const Particle &p, double vx, double vy, double vz, int gid)
{
	double2 *q = reinterpret_cast<double2 *>(e);
	long long w = ((long long)(unsigned long long)p.cell << 32) | (unsigned)p.type;
	q[0] = make_double2(p.x, p.y);
	q[1] = make_double2(p.z, __longlong_as_double(w));
	q[2] = make_double2(vx, vy);
	q[3] = make_double2(vz, __longlong_as_double((long long)(unsigned)gid));
}

// One owned particle (or none: own == false) of the exchange, called by all 32 lanes of a warp: p carries the new position
// and cell tag; (vx, vy, vz) and (ux, uy, uz) travel only with migrants.  Marks a migrant as a ghost in gid[s].
__device__ __forceinline__ bool slab_pack_one(bool own, int s, const Particle &p, int gi, double vx, double vy, double vz, bool has_unw,
                                              double ux, double uy, double uz, int *gid, const Geom &g, const SlabComm &c, int seq, int *errflag)
{
	int dir = -1;          // 0: entry for the left neighbour, 1: for the right one
	bool migrant = false, wrote = false;
	if (own) {
		int cx, cy, cz;
		unpack_cell(p.cell, cx, cy, cz);
		int W = g.col_hi - g.col_lo;
		int rel = cx - g.col_lo;
		if (rel < 0) rel += g.nc[0];
		else if (rel >= g.nc[0]) rel -= g.nc[0];
		if (rel < W) {
			if (rel < g.halo) dir = 0;
			else if (rel >= W - g.halo) dir = 1;
		} else if (rel < W + g.halo) {
			dir = 1; migrant = true;
		} else if (rel >= g.nc[0] - g.halo) {
			dir = 0; migrant = true;
		} else {
			atomicOr(errflag, ERR_SLAB_MIGRATION);   // moved further than the halo in one step
		}
		if (migrant) gid[s] = gi | GID_GHOST;
	}
#pragma unroll
	for (int d = 0; d < 2; d++) {
		unsigned m = __ballot_sync(0xffffffffu, dir == d);
		if (m == 0) continue;
		int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
		int basei = 0;
		if (lane == leader) basei = atomicAdd(c.counters + d, __popc(m));
		basei = __shfl_sync(0xffffffffu, basei, leader);
		if (dir == d) {
			int idx = basei + __popc(m & ((1u << lane) - 1u));
			if (idx >= c.capmsg) {
				atomicOr(errflag, ERR_SLAB_MSG_CAP);
			} else {
				char *buf = (c.pull ? c.recv[d] : c.send[d]) + (size_t)(seq & 1) * c.parity_stride;
				SlabMsgEntry *e = reinterpret_cast<SlabMsgEntry *>(buf + sizeof(SlabMsgHeader)) + idx;
				st_entry(e, p, migrant ? vx : 0.0, migrant ? vy : 0.0, migrant ? vz : 0.0, migrant ? gi : (gi | GID_GHOST));
				if (migrant && has_unw) {
					double2 *q = reinterpret_cast<double2 *>(e);
					q[4] = make_double2(ux, uy);
					q[5] = make_double2(uz, 0.0);
				}
				wrote = true;
			}
		}
	}
	return wrote;
}

// publish (all threads of every block of the grid that packed): the threads that wrote an entry fence it at system scope
// (only they: a system fence waits for everything the thread has in flight, and in the step seam that is every thread's
// position / velocity stores -- fencing all 10^6 threads cost more than the pack pass it replaced), the last block to
// arrive writes the two headers
__device__ __forceinline__ void slab_publish(const SlabComm &c, int seq, bool wrote)
{
	// pull: the entries are in this device's own memory: a device-scope fence orders them before the arrival below, the last
	// block observes every arrival and then releases the headers at system scope (release is cumulative: what the releasing
	// thread has observed is visible to whoever acquires the header).  push: the entries crossed the link, see SlabComm.
	// SMD_SLAB_FENCE_GPU=1 (experiment, push): writers fence at device scope only and thread 0 releases the block's arrival with a
	// device-scope fence of its own; by the PTX model's cumulativity the last block's system-scope release then covers every
	// entry (writer -> barrier -> release / acquire chain on the counter -> system fence -> header).  2 x B200: seam 98 -> 86 us,
	// 860.4 -> 856.4 us per step, trajectories bit-identical over 2 100 steps.  NOT the default: it leans on a system fence of one SM
	// completing peer writes that other SMs still have in flight, which the model promises and no test here can prove.
	if (wrote) { if (c.pull || c.gpu_fence) __threadfence(); else __threadfence_system(); }
	__syncthreads();
	if (threadIdx.x == 0) {
		if (c.gpu_fence) __threadfence();
		int t = atomicAdd(c.counters + 2, 1);
		if (t == (int)gridDim.x - 1) {
			__threadfence_system();
			int n0 = min(atomicExch(c.counters + 0, 0), c.capmsg), n1 = min(atomicExch(c.counters + 1, 0), c.capmsg);
			c.counters[2] = 0;
			__threadfence_system();
			volatile int2 *h0 = reinterpret_cast<volatile int2 *>((c.pull ? c.recv[0] : c.send[0]) + (size_t)(seq & 1) * c.parity_stride);
			volatile int2 *h1 = reinterpret_cast<volatile int2 *>((c.pull ? c.recv[1] : c.send[1]) + (size_t)(seq & 1) * c.parity_stride);
			int2 v0 = make_int2(n0, seq), v1 = make_int2(n1, seq);
			asm volatile("st.release.sys.global.v2.s32 [%0], {%1, %2};" ::"l"(h0), "r"(v0.x), "r"(v0.y) : "memory");
			asm volatile("st.release.sys.global.v2.s32 [%0], {%1, %2};" ::"l"(h1), "r"(v1.x), "r"(v1.y) : "memory");
		}
	}
}

__global__ void __launch_bounds__(TPB) k_slab_pack(Cnt cnt, int cap, Particle *pos, const double *vel, const double *unw, int *gid, Geom g,
                                                   SlabComm c, int seq, int *errflag, int *slot_of)
{
	const int N = cnt.get();
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	bool own = false;
	Particle p;
	p.x = p.y = p.z = 0; p.type = 0; p.cell = 0;
	int gi = 0;
	double vx = 0, vy = 0, vz = 0, ux = 0, uy = 0, uz = 0;
	if (s < N) {
		gi = gid[s];
		if (gi & GID_GHOST) {
			pos[s].cell = CELL_DEAD;
			slot_of[gi & GID_MASK] = -1;   // keeps the index table exact: whoever is still here is re-entered by the build
		} else {
			own = true;
			p = load_particle(pos + s);
			vx = vel[s]; vy = vel[cap + s]; vz = vel[2 * cap + s];
			if (unw) { ux = unw[s]; uy = unw[cap + s]; uz = unw[2 * cap + s]; }
		}
	}
	const bool wrote = slab_pack_one(own, s, p, gi, vx, vy, vz, unw != nullptr, ux, uy, uz, gid, g, c, seq, errflag);
	slab_publish(c, seq, wrote);
}

__global__ void __launch_bounds__(256) k_slab_unpack(Cnt cnt, int *dNext, int cap, Particle *pos, double *vel, double *unw, int *gid,
                                                     SlabComm c, int seq, int *errflag, long long spin_limit, Geom g, BinArgs bin)
{
	// (not pdl_prologue(): this kernel SPINS on its neighbours' headers.  Its dependents -- the persistent grid of k_scan is
	// next -- are released only once the messages are in: waiting blocks of several ranks that share one device (the
	// single-process test harness) could otherwise fill every SM while the kernel that sends the message finds no room)
	asm volatile("griddepcontrol.wait;" ::: "memory");
	__shared__ int n_s[2];
	if (threadIdx.x < 2) {
		const char *buf = (c.pull ? c.send[threadIdx.x] : c.recv[threadIdx.x]) + (size_t)(seq & 1) * c.parity_stride;
		int n = 0, sq = seq - 1;
		long long t0 = clock64();
		while (true) {
			asm volatile("ld.acquire.sys.global.v2.s32 {%0, %1}, [%2];" : "=r"(n), "=r"(sq) : "l"(buf) : "memory");
			if (sq == seq) break;
			if (clock64() - t0 > spin_limit) { atomicOr(errflag, ERR_SLAB_TIMEOUT); n = 0; break; }
			__nanosleep(200);
		}
		n_s[threadIdx.x] = n;
	}
	__syncthreads();
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	const int N0 = cnt.get(), nL = n_s[0], nR = n_s[1], total = nL + nR;
	for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
		int side = e >= nL, k = e - (side ? nL : 0);
		const char *buf = (c.pull ? c.send[side] : c.recv[side]) + (size_t)(seq & 1) * c.parity_stride;
		const double2 *q = reinterpret_cast<const double2 *>(reinterpret_cast<const SlabMsgEntry *>(buf + sizeof(SlabMsgHeader)) + k);
		int slot = N0 + e;
		if (slot >= cap) { atomicOr(errflag, ERR_SLAB_CAPACITY); continue; }
		// (pull: the neighbour's memory, read through the link: L2-only loads, nothing of an earlier exchange can sit in an L1)
		double2 a = __ldcg(q), b = __ldcg(q + 1), v0 = __ldcg(q + 2), v1 = __ldcg(q + 3);
		double2 *o = reinterpret_cast<double2 *>(pos + slot);
		o[0] = a; o[1] = b;
		vel[slot] = v0.x; vel[cap + slot] = v0.y; vel[2 * cap + slot] = v1.x;
		gid[slot] = (int)(unsigned)(__double_as_longlong(v1.y) & 0xffffffffll);
		if (unw) {
			double2 u0 = __ldcg(q + 4), u1 = __ldcg(q + 5);
			unw[slot] = u0.x; unw[cap + slot] = u0.y; unw[2 * cap + slot] = u1.x;
		}
		if (bin.count) {   // the histogram of the build that follows: the seam binned what stayed, this bins what arrived
			// (the record carries the sender's cell tag: the reference grid is the same on every rank)
			int cx, cy, cz;
			unpack_cell((unsigned)((unsigned long long)__double_as_longlong(b.y) >> 32), cx, cy, cz);
			const int lx = win_x(cx, bin.win[WIN_ORG], g.nc[0]);
			int local = -1;
			if (lx >= bin.win[WIN_DIM]) atomicOr(errflag, ERR_SLAB_MIGRATION);
			else {
				int sub = 0;
				if (g.xs > 1) sub = min(max((int)((a.x - (double)cx * g.cs[0]) * g.finv), 0), g.xs - 1);
				local = (lx * g.xs + sub) + bin.win[WIN_FD0] * ((cy - bin.win[WIN_ORG + 1]) + bin.win[WIN_DIM + 1] * (cz - bin.win[WIN_ORG + 2]));
				atomicAdd(bin.count + local, 1);
			}
			bin.cellOfSlot[slot] = local;
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) *dNext = min(N0 + total, cap);
}

// owned particles of a slab, compacted (order arbitrary): global index, record, velocity
__global__ void __launch_bounds__(TPB) k_slab_export(Cnt cnt, int cap, const Particle *pos, const double *vel, const double *acc,
                                                     const int *gid, int *counter, int *out_gid, double *out_xyz, double *out_vel,
                                                     double *out_acc, int *out_type)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	bool own = s < cnt.get() && !(gid[s] & GID_GHOST) && pos[s].cell != CELL_DEAD;
	unsigned m = __ballot_sync(0xffffffffu, own);
	if (m == 0) return;
	int lane = threadIdx.x & 31, leader = __ffs(m) - 1, b = 0;
	if (lane == leader) b = atomicAdd(counter, __popc(m));
	b = __shfl_sync(0xffffffffu, b, leader);
	if (!own) return;
	int k = b + __popc(m & ((1u << lane) - 1u));
	Particle p = load_particle(pos + s);
	out_gid[k] = gid[s];
	out_xyz[3 * k] = p.x; out_xyz[3 * k + 1] = p.y; out_xyz[3 * k + 2] = p.z;
	out_type[k] = p.type;
	out_vel[3 * k] = vel[s]; out_vel[3 * k + 1] = vel[cap + s]; out_vel[3 * k + 2] = vel[2 * cap + s];
	out_acc[3 * k] = acc[s]; out_acc[3 * k + 1] = acc[cap + s]; out_acc[3 * k + 2] = acc[2 * cap + s];
}

// ------------------------------------------------------------------------------------------------ fused step seam (kernel; chain_gather is defined ahead of the pair kernel)
#ifndef SMD_KICK_BLOCKS
#define SMD_KICK_BLOCKS 5   // 96 registers with a few spilled words beat 122 registers at 4 blocks (22.9 vs 25.1 us on C2)
#endif
// Continuum-sphere particles (BEAD and NANOCORE molecules) are divided by their "mass" 4 pi R^2 between the force
// evaluation and the kicks: BEAD beads before Verlet::second (MD.cpp:480-494) AND again before the next Verlet::first
// (:340-355, quirk Q3), NANOCORE beads only before Verlet::second (:495-508).  A handful of particles: a by-value list.
constexpr int MAX_FUSED_BEADS = 8;
struct BeadSet { int n; int id[MAX_FUSED_BEADS]; int twice[MAX_FUSED_BEADS]; double mass[MAX_FUSED_BEADS]; };

template <bool LAST>
__global__ void __launch_bounds__(TPB, SMD_KICK_BLOCKS) k_chain_kick(Cnt cnt, int cap, const Particle *__restrict__ pos_in, Particle *__restrict__ pos_out,
                                                    double *vel, double *acc, double *unw, const int *__restrict__ gid,
                                                    const int *__restrict__ slot_of, Geom g, ChainSet cs, double dt, int *bbox, int *errflag,
                                                    BeadSet bs, int slot0, SlabComm comm, int seq, int *gid_w, const int *done, int epoch,
                                                    BinArgs bin, int done_per = 1)
{
	SMD_TL(4);
	// done != nullptr: launched as a programmatic dependent of the pair kernel -- this block may be resident while the
	// tail of that grid is still running and only needs the accelerations of its own 128 slots, which pair block `blk`
	// signals with done[blk] = epoch (everything else it reads was final before the pair kernel started)
	// without completion words (systems with list molecules or fields: their kernels ran in between) the plain dependency
	if (!done) asm volatile("griddepcontrol.wait;" ::: "memory");
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the build that follows waits for this grid (pdl_prologue)
	if (done) {
		const int blk = slot0 / TPB + (int)blockIdx.x;
		if (blk * TPB < cnt.get()) {
			// done_per pair blocks cover this block's slots (4 when the pair kernel runs 32 particles per block)
			const int sub = TPB / done_per;
			if ((int)threadIdx.x < done_per && blk * TPB + (int)threadIdx.x * sub < cnt.get()) {
				int v;
				do {
					asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(done + blk * done_per + threadIdx.x) : "memory");
				} while (v != epoch);
			}
			__syncthreads();
		}
	}
	// seq > 0 (slab mode, not LAST): this kernel is also the SEND side of the step's halo / migration exchange -- every
	// owned particle is packed as soon as it has its new position (slab_pack_one), straight into the neighbour's receive
	// buffer through peer memory, and the last block publishes the headers: no separate pass over the slots
	const int N = cnt.get();
	int s = slot0 + blockIdx.x * blockDim.x + threadIdx.x;   // slot0 != 0: one chunk of the step pipeline (smd_step)
	bool valid = s < N;
	Particle p;
	int gi = 0;
	if (valid) { p = load_particle(pos_in + s); gi = gid[s]; }
	bool live = valid && !(g.slab && (gi & GID_GHOST));
	double mvx = 0, mvy = 0, mvz = 0, mux = 0, muy = 0, muz = 0;   // what a migrant takes along (slab mode)
	if (live) {
		V3 A;
		if (!chain_gather(gi & GID_MASK, p, N, pos_in, gid, slot_of, g, cs, A)) atomicOr(errflag, ERR_SLAB_MISSING);
		double ax = acc[s] + A.x, ay = acc[cap + s] + A.y, az = acc[2 * cap + s] + A.z;   // Blob::doChainForce: a += chain terms
		double bx = ax, by = ay, bz = az;   // the acceleration the NEXT Verlet::first sees
		for (int b = 0; b < bs.n; b++)
			if ((gi & GID_MASK) == bs.id[b]) {
				ax /= bs.mass[b]; ay /= bs.mass[b]; az /= bs.mass[b];                       // MD.cpp:480-508
				bx = ax; by = ay; bz = az;
				if (bs.twice[b]) { bx /= bs.mass[b]; by /= bs.mass[b]; bz /= bs.mass[b]; }  // MD.cpp:340-355
			}
		if (LAST) { acc[s] = ax; acc[cap + s] = ay; acc[2 * cap + s] = az; }
		if (p.type != 0) {
			double h = 0.5 * dt;
			double vx = vel[s], vy = vel[cap + s], vz = vel[2 * cap + s];
			vx += (ax * h); vy += (ay * h); vz += (az * h);                                 // Verlet::second of this step
			if (!LAST) {
				vx += (bx * h); vy += (by * h); vz += (bz * h);                             // Verlet::first of the next one
				p.x += vx * dt; p.y += vy * dt; p.z += vz * dt;
				if (unw) {
					mux = unw[s] + vx * dt; muy = unw[cap + s] + vy * dt; muz = unw[2 * cap + s] + vz * dt;
					unw[s] = mux; unw[cap + s] = muy; unw[2 * cap + s] = muz;
				}
			}
			vel[s] = vx; vel[cap + s] = vy; vel[2 * cap + s] = vz;
			mvx = vx; mvy = vy; mvz = vz;
		} else if (!LAST && seq > 0) {   // type 0 never moves, but it is exchanged like everybody else
			mvx = vel[s]; mvy = vel[cap + s]; mvz = vel[2 * cap + s];
			if (unw) { mux = unw[s]; muy = unw[cap + s]; muz = unw[2 * cap + s]; }
		}
		if (!LAST) {
			if (p.x > g.box[0]) p.x -= g.box[0];
			if (p.x < 0) p.x += g.box[0];
			if (p.y > g.box[1]) p.y -= g.box[1];
			if (p.y < 0) p.y += g.box[1];
			if (p.z > g.box[2]) p.z -= g.box[2];
			if (p.z < 0) p.z += g.box[2];
		}
	}
	if (LAST) return;
	tag_cell(p, g, bbox, errflag, live);
	if (valid) {
		if (!live) p.cell = CELL_DEAD;   // slab ghost: replaced by the exchange (its index-table entry is given back by k_bin)
		store_particle(pos_out + s, p);
	}
	if (bin.count) {   // the histogram of the build that follows
		const int local = bin_particle(p, live, g, bin, errflag);
		if (valid) bin.cellOfSlot[s] = local;
	}
	if (seq > 0) {
		const bool wrote = slab_pack_one(live, s, p, gi, mvx, mvy, mvz, unw != nullptr, mux, muy, muz, gid_w, g, comm, seq, errflag);
		slab_publish(comm, seq, wrote);
	}
}

// explicit BOND list (system.h:1880-1934, :2717-2747, :3398-3435); one thread per bond, FP64 atomics for the force
// Slab mode (records carry global indices; every rank walks the whole list): a record is evaluated by every rank that OWNS
// one of its members -- the others must then be inside its halo, else ERR_SLAB_MISSING --, forces go to owned members only,
// and the energy terms are counted by the rank that owns the FIRST member, so every record is counted exactly once.
__device__ __forceinline__ bool slab_owned(int s, const int *__restrict__ gid) { return s >= 0 && !(gid[s] & GID_GHOST); }

template <int MODE>
__global__ void __launch_bounds__(TPB) k_bond(int nb, int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of, Geom g,
                                              const int *__restrict__ ij, double r0, double kk, double *acc, double *partials,
                                              double sx, double sy, double sz, const int *__restrict__ gid, int *errflag)
{
	pdl_prologue();
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	double usum = 0;
	bool o1 = true, o2 = true, go = b < nb;
	int s1 = 0, s2 = 0;
	if (go) {
		s1 = slot_of[ij[2 * b]]; s2 = slot_of[ij[2 * b + 1]];
		if (g.slab) {
			o1 = slab_owned(s1, gid); o2 = slab_owned(s2, gid);
			go = MODE == 0 ? (o1 || o2) : o1;
			if (go && (s1 < 0 || s2 < 0)) { atomicOr(errflag, ERR_SLAB_MISSING); go = false; }
		}
	}
	if (go) {
		V3 d = diff_mi(load_particle(pos + s1), load_particle(pos + s2), g);
		if (MODE == 0) {
			V3 f = harmonic_f(d, r0, kk);
			if (o1) { atomicAdd(acc + s1, f.x); atomicAdd(acc + cap + s1, f.y); atomicAdd(acc + 2 * cap + s1, f.z); }
			if (o2) { atomicAdd(acc + s2, -f.x); atomicAdd(acc + cap + s2, -f.y); atomicAdd(acc + 2 * cap + s2, -f.z); }
		} else if (MODE == 1) {
			usum = harmonic_p(d, r0, kk);
		} else {
			double uo = harmonic_p(d, r0, kk);
			usum = uo - harmonic_p(scaled(d, sx, sy, sz), r0, kk);
		}
	}
	if (MODE != 0) {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
}

// BALL (system.h:1936-1971, :2751-2781, :3439-3474; harmonicHalfF / harmonicHalfP, MD.h:1164-1222): a harmonic wall
// around a centre particle that acts only INSIDE r0.  One thread per record {centre, j}; the centre's share is
// block-reduced and added with one atomic per block.  Quirk reproduced: the force routine takes the centre of
// EVERY record from record 0 (`first=bond[0].s[0]`, system.h:1943), the two energy routines read each record's own.
template <int MODE>
__global__ void __launch_bounds__(TPB) k_ball(int nb, int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of, Geom g,
                                              const int *__restrict__ cj, double r0, double kk, double *acc, double *partials,
                                              double sx, double sy, double sz)
{
	pdl_prologue();
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	double usum = 0, fx = 0, fy = 0, fz = 0;
	const int s1f = (MODE == 0) ? slot_of[cj[0]] : 0;
	auto half_p = [&](V3 d) {
		double dr = d.x * d.x + d.y * d.y + d.z * d.z;
		double u = 0;
		if (dr < r0 * r0) {
			dr = sqrt(dr);
			u = dr - r0;
			u = 0.5 * kk * u * u;
		}
		return u;
	};
	if (b < nb) {
		int s1 = (MODE == 0) ? s1f : slot_of[cj[2 * b]], s2 = slot_of[cj[2 * b + 1]];
		V3 d = diff_mi(load_particle(pos + s1), load_particle(pos + s2), g);
		if (MODE == 0) {
			double dr = d.x * d.x + d.y * d.y + d.z * d.z;
			if (dr < r0 * r0) {
				dr = sqrt(dr);
				double m = dr - r0;
				m = -m * kk / dr;
				fx = d.x * m; fy = d.y * m; fz = d.z * m;
				atomicAdd(acc + s2, -fx); atomicAdd(acc + cap + s2, -fy); atomicAdd(acc + 2 * cap + s2, -fz);
			}
		} else if (MODE == 1) {
			usum = half_p(d);
		} else {
			double uo = half_p(d);
			usum = uo - half_p(scaled(d, sx, sy, sz));
		}
	}
	if (MODE == 0) {
		fx = block_sum(fx); fy = block_sum(fy); fz = block_sum(fz);
		if (threadIdx.x == 0) { atomicAdd(acc + s1f, fx); atomicAdd(acc + cap + s1f, fy); atomicAdd(acc + 2 * cap + s1f, fz); }
	} else {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
}

// explicit BEND list (system.h:1975-2040, :2781-2822, :3476-3527)
template <int MODE>
__global__ void __launch_bounds__(TPB) k_bend(int nb, int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of, Geom g,
                                              const int *__restrict__ ijk, double c0, double kk, double *acc, double *partials,
                                              double sx, double sy, double sz, const int *__restrict__ gid, int *errflag)
{
	pdl_prologue();
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	double usum = 0;
	bool o1 = true, o2 = true, o3 = true, go = b < nb;
	int s1 = 0, s2 = 0, s3 = 0;
	if (go) {
		s1 = slot_of[ijk[3 * b]]; s2 = slot_of[ijk[3 * b + 1]]; s3 = slot_of[ijk[3 * b + 2]];
		if (g.slab) {   // (see k_bond)
			o1 = slab_owned(s1, gid); o2 = slab_owned(s2, gid); o3 = slab_owned(s3, gid);
			go = MODE == 0 ? (o1 || o2 || o3) : o1;
			if (go && (s1 < 0 || s2 < 0 || s3 < 0)) { atomicOr(errflag, ERR_SLAB_MISSING); go = false; }
		}
	}
	if (go) {
		Particle p2 = load_particle(pos + s2);
		V3 da = diff_mi(load_particle(pos + s1), p2, g), db = diff_mi(p2, load_particle(pos + s3), g);
		if (MODE == 0) {
			V3 fa, fb;
			bend_f(da, db, c0, kk, fa, fb);
			if (o1) { atomicAdd(acc + s1, fa.x); atomicAdd(acc + cap + s1, fa.y); atomicAdd(acc + 2 * cap + s1, fa.z); }
			if (o2) { atomicAdd(acc + s2, fb.x - fa.x); atomicAdd(acc + cap + s2, fb.y - fa.y); atomicAdd(acc + 2 * cap + s2, fb.z - fa.z); }
			if (o3) { atomicAdd(acc + s3, -fb.x); atomicAdd(acc + cap + s3, -fb.y); atomicAdd(acc + 2 * cap + s3, -fb.z); }
		} else if (MODE == 1) {
			usum = bend_p(da, db, c0, kk);
		} else {
			double uo = bend_p(da, db, c0, kk);
			usum = uo - bend_p(scaled(da, sx, sy, sz), scaled(db, sx, sy, sz), c0, kk);
		}
	}
	if (MODE != 0) {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
}

// ------------------------------------------------------------------------------------------------ BEAD molecule
// continuum-sphere bead terms, MD.h:165-382

__device__ __forceinline__ double bead_mag(double dr2, const double *C, double cut2)
{
	double mag = 0;
	double rminD = C[4] + C[5];
	double rminD2 = rminD * rminD;
	if (rminD2 <= dr2 && dr2 < cut2) {
		double dr = sqrt(dr2);
		double E = C[0] - dr, E2 = E * E, E3 = E * E2;
		mag = 2.0 * (0.4 * E3 + C[1] * E2 + C[2] * E) / (dr);
		mag += 4.0 * E2 + 8.0 * C[1] * E + 6.0 * C[2];
		mag *= (E2 * C[3]) / (dr * dr);
	} else if (dr2 < rminD2) {
		double dr = sqrt(dr2);
		double B = rminD - dr, B2 = B * B, B3 = B2 * B, B4 = B2 * B2;
		mag = (B4 * C[6] + B3 * C[7] + B2 * C[8] + B * C[9] + C[10]) / (dr);
		mag += (4.0 * B3 * C[6] + 3.0 * B2 * C[7] + 2.0 * B * C[8] + C[9]);
		mag /= dr * dr;
	}
	return mag;
}

__device__ __forceinline__ double bead_pot(double dr2, const double *C, double cut2)
{
	double u = 0;
	double rminD = C[4] + C[5];
	double rminD2 = rminD * rminD;
	if (rminD2 <= dr2 && dr2 < cut2) {
		double dr = sqrt(dr2);
		double E = C[0] - dr, E2 = E * E, E3 = E * E2;
		u = 2.0 * E3 * (0.4 * E2 + C[1] * E + C[2]) * C[3] / (dr);
	} else if (dr2 < rminD2) {
		double dr = sqrt(dr2);
		double B = rminD - dr, B2 = B * B, B3 = B2 * B, B4 = B2 * B2;
		u = (B4 * C[6] + B3 * C[7] + B2 * C[8] + B * C[9] + C[10]) / (dr);
	}
	return u;
}

__device__ __forceinline__ double beadbead_mag(double dr2, const double *C, double cut2)
{
	const double *K = C + 11;   // BEADBEADOFFSET, MD.h:51
	double mag = 0;
	double rminD = K[0], rminD2 = rminD * rminD;
	if (rminD2 <= dr2 && dr2 < cut2) {
		double dr = sqrt(dr2);
		double E = K[1] - dr, E2 = E * E, E3 = E * E2;
		mag = E3 * (6.0 * K[2] * E2 + 5.0 * K[3] * E + 4.0 * K[4] + (K[2] * E3 + K[3] * E2 + K[4] * E) / dr) / (dr * dr);
	} else if (dr2 < rminD2) {
		double dr = sqrt(dr2);
		double B = rminD - dr, B2 = B * B, B3 = B2 * B, B4 = B2 * B2, B5 = B2 * B3;
		mag = 5.0 * B4 * K[5] + 4.0 * B3 * K[6] + 3.0 * B2 * K[7] + 2.0 * B * K[8] + K[9];
		mag += (B5 * K[5] + B4 * K[6] + B3 * K[7] + B2 * K[8] + B * K[9] + K[10]) / (dr);
		mag /= dr * dr;
	}
	return mag;
}

__device__ __forceinline__ double beadbead_pot(double dr2, const double *C, double cut2)
{
	const double *K = C + 11;
	double u = 0;
	double rminD = K[0], rminD2 = rminD * rminD;
	if (rminD2 <= dr2 && dr2 < cut2) {
		double dr = sqrt(dr2);
		double E = K[1] - dr, E2 = E * E, E4 = E2 * E2;
		u = E4 * (K[2] * E2 + K[3] * E + K[4]) / dr;
	} else if (dr2 < rminD2) {
		double dr = sqrt(dr2);
		double B = rminD - dr, B2 = B * B, B3 = B2 * B, B4 = B2 * B2, B5 = B2 * B3;
		u = (B5 * K[5] + B4 * K[6] + B3 * K[7] + B2 * K[8] + B * K[9] + K[10]) / (dr);
	}
	return u;
}

// bead-bead pairs of one BEAD molecule: j over own beads, k > j over the assembled list (system.h:2074-2098).
// A handful of pairs: one block, thread per (j,k).
template <int MODE>
__global__ void __launch_bounds__(TPB) k_beadbead(int nOwn, int nAll, int cap, const Particle *__restrict__ pos,
                                                  const int *__restrict__ slot_of, Geom g, int nT, const int *__restrict__ beads,
                                                  const double *__restrict__ C, double R, double *acc, double *partials,
                                                  double sx, double sy, double sz)
{
	pdl_prologue();
	double usum = 0;
	for (int t = threadIdx.x; t < nOwn * nAll; t += blockDim.x) {
		int j = t / nAll, k = t % nAll;
		if (k <= j) continue;
		int s1 = slot_of[beads[j]], s2 = slot_of[beads[k]];
		Particle p1 = load_particle(pos + s1), p2 = load_particle(pos + s2);
		V3 d = diff_mi(p1, p2, g);
		double cut = R + R + 2.0;
		cut *= cut;
		const double *Cr = C + 22 * (p1.type * nT + p2.type);
		double dr2 = d.x * d.x + d.y * d.y + d.z * d.z;
		if (MODE == 0) {
			double m = beadbead_mag(dr2, Cr, cut);
			double fx = d.x * m, fy = d.y * m, fz = d.z * m;
			atomicAdd(acc + s1, fx); atomicAdd(acc + cap + s1, fy); atomicAdd(acc + 2 * cap + s1, fz);
			atomicAdd(acc + s2, -fx); atomicAdd(acc + cap + s2, -fy); atomicAdd(acc + 2 * cap + s2, -fz);
		} else if (MODE == 1) {
			usum += beadbead_pot(dr2, Cr, cut);
		} else {
			double uo = beadbead_pot(dr2, Cr, cut);
			V3 e = scaled(d, sx, sy, sz);
			usum += (uo - beadbead_pot(e.x * e.x + e.y * e.y + e.z * e.z, Cr, cut));
		}
	}
	if (MODE != 0) {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[0] = usum;
	}
}

// bead - particle terms THROUGH THE CELL GRID: one block per bead; only the particles in the cells within R + rc of the
// bead are visited (the reference's brute-force branch system.h:2165-2210 tests every particle against the same cutoff, its
// hash-cell branch :2105-2164 -- more than 20 beads -- walks cells of 2 (R + rc); all three visit the same pairs).  The
// cube of cells around the bead is cut into (y,z) rows, a row into at most two x segments (periodic wrap), a segment is
// one contiguous slot range of the cell-sorted order, clipped to the occupied window (cells outside it are empty).
// excl_all: the force excludes every own bead when nOwn <= 20 and only the bead itself otherwise; potential / dPotential
// always exclude only the bead itself (Q7).
// nano: NANOCORE molecules (doNanoCoreForce system.h:2215-2332, doNanoCorePotential :3028-3075, doNanoCoreDPotential
// :3708-3760) are the same term with ONE constants row per bead (C + 22 j, cutoff from that row) whatever the particle's
// type, only the bead itself excluded, and no bead-bead term.
// Force on a particle: one FP64 atomic per component (several beads may touch it); reaction on the bead: summed in
// registers over the whole cube, one block reduction and one atomic per bead.
constexpr int BEAD_SEGS = 1024;
template <int MODE>
__global__ void __launch_bounds__(TPB) k_bead(int nOwn, int cap, const Particle *__restrict__ pos, const int *__restrict__ gid,
                                              const int *__restrict__ slot_of, const int *__restrict__ start, const int *__restrict__ win,
                                              Geom g, int nT, const int *__restrict__ beads, const double *__restrict__ C, int excl_all,
                                              int nano, double *acc, double *partials, double sx, double sy, double sz, int nparts)
{
	pdl_prologue();
	__shared__ int seg_b[BEAD_SEGS], seg_e[BEAD_SEGS];
	// nparts blocks per bead: block (j, part) takes the rows part, part + nparts, ... of the bead's cube, so that ONE large
	// bead (C3: a sphere of radius 5.88 against 3 000 particles) is not one block's serial loop; grid = nOwn * nparts
	const int j = blockIdx.x / nparts, part = blockIdx.x % nparts;
	const int bs = slot_of[beads[j]];
	const Particle pb = load_particle(pos + bs);
	const double cutR = nano ? C[22 * j] : C[0];   // R + rc (BEAD: row 0, system.h:2100-2101; NANOCORE: the bead's own row)
	const double cut2 = cutR * cutR;
	// MODE 2: a pair outside the cutoff may be inside it after the proposed scaling and then still has a term
	const double reach = MODE == 2 ? cutR * fmax(1.0, 1.0 / fmin(sx, fmin(sy, sz))) * (1.0 + 1e-9) : cutR;
	const double reach2 = reach * reach;
	int c[3];
	{ int cx, cy, cz; unpack_cell(pb.cell, cx, cy, cz); c[0] = cx; c[1] = cy; c[2] = cz; }
	// cell layers around the bead's cell that can hold a particle within the cutoff, and cells per axis to visit
	int k[3], cnt[3];
	for (int d = 0; d < 3; d++) {
		k[d] = (int)(reach / g.cs[d]) + 1;
		cnt[d] = min(2 * k[d] + 1, g.nc[d]);   // a cube wider than the box: every cell once
	}
	const int w[3] = {win[WIN_ORG], win[WIN_ORG + 1], win[WIN_ORG + 2]};
	const int dm[3] = {win[WIN_DIM], win[WIN_DIM + 1], win[WIN_DIM + 2]};
	const int fd0 = win[WIN_FD0], xs = g.xs;
	auto axis_cell = [&](int d, int i) {   // i-th visited cell along axis d
		if (cnt[d] == g.nc[d]) return i;
		int v = c[d] - k[d] + i;
		if (v < 0) v += g.nc[d];
		if (v >= g.nc[d]) v -= g.nc[d];
		return v;
	};
	double usum = 0, rx = 0, ry = 0, rz = 0;
	const int nrows = cnt[1] * cnt[2];
	const int nmine = part < nrows ? (nrows - part + nparts - 1) / nparts : 0;   // rows of this block
	for (int r0 = 0; r0 < nmine; r0 += BEAD_SEGS / 2) {
		// ---- the slot ranges of up to BEAD_SEGS / 2 rows, two x segments each
		const int nq = min(BEAD_SEGS / 2, nmine - r0);
		__syncthreads();
		for (int q = threadIdx.x; q < nq; q += blockDim.x) {
			int b0 = 0, e0 = 0, b1 = 0, e1 = 0;
			const int r = part + nparts * (r0 + q);
			if (r < nrows) {
				const int ly = axis_cell(1, r % cnt[1]) - w[1], lz = axis_cell(2, r / cnt[1]) - w[2];
				if (ly >= 0 && ly < dm[1] && lz >= 0 && lz < dm[2]) {
					const int rowbase = fd0 * (ly + dm[1] * lz);
					auto range = [&](int xlo, int xhi, int &b, int &e) {   // cells xlo .. xhi (no wrap inside), clipped to the window
						xlo = max(xlo - w[0], 0); xhi = min(xhi - w[0], dm[0] - 1);
						if (xlo <= xhi) { b = start[rowbase + xlo * xs]; e = start[rowbase + (xhi + 1) * xs]; }
					};
					if (cnt[0] == g.nc[0]) range(0, g.nc[0] - 1, b0, e0);
					else {
						const int lo = c[0] - k[0], hi = c[0] + k[0];
						if (lo < 0) { range(lo + g.nc[0], g.nc[0] - 1, b0, e0); range(0, hi, b1, e1); }
						else if (hi >= g.nc[0]) { range(lo, g.nc[0] - 1, b0, e0); range(0, hi - g.nc[0], b1, e1); }
						else range(lo, hi, b0, e0);
					}
				}
			}
			seg_b[2 * q] = b0; seg_e[2 * q] = e0; seg_b[2 * q + 1] = b1; seg_e[2 * q + 1] = e1;
		}
		__syncthreads();
		// ---- the particles of those ranges
		for (int sgm = 0; sgm < 2 * nq; sgm++) {
			const int e = seg_e[sgm];
			for (int s = seg_b[sgm] + threadIdx.x; s < e; s += blockDim.x) {
				if (s == bs) continue;
				const Particle p = load_particle(pos + s);
				if (excl_all) {
					const int id = gid[s] & GID_MASK;
					bool own = false;
					for (int q = 0; q < nOwn; q++) own |= (beads[q] == id);
					if (own) continue;
				}
				V3 d = diff_mi(pb, p, g);
				const double dr2 = d.x * d.x + d.y * d.y + d.z * d.z;
				if (!(dr2 < reach2)) continue;
				const double *Cr = nano ? C + 22 * j : C + 22 * (pb.type * nT + p.type);
				if (MODE == 0) {
					const double m = bead_mag(dr2, Cr, cut2);
					if (m != 0) {
						const double fx = d.x * m, fy = d.y * m, fz = d.z * m;
						atomicAdd(acc + s, -fx); atomicAdd(acc + cap + s, -fy); atomicAdd(acc + 2 * cap + s, -fz);
						rx += fx; ry += fy; rz += fz;
					}
				} else if (MODE == 1) {
					usum += bead_pot(dr2, Cr, cut2);
				} else {
					const double uo = bead_pot(dr2, Cr, cut2);
					V3 e2 = scaled(d, sx, sy, sz);
					usum += (uo - bead_pot(e2.x * e2.x + e2.y * e2.y + e2.z * e2.z, Cr, cut2));
				}
			}
		}
	}
	if (MODE == 0) {
		rx = block_sum(rx); ry = block_sum(ry); rz = block_sum(rz);
		if (threadIdx.x == 0 && (rx != 0 || ry != 0 || rz != 0)) { atomicAdd(acc + bs, rx); atomicAdd(acc + cap + bs, ry); atomicAdd(acc + 2 * cap + bs, rz); }
	} else {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
}

// ------------------------------------------------------------------------------------------------ external fields
// The remaining molecule kinds of MD.cpp's switch (MD.cpp:414-478).  One thread per record (or per bond of
// a block), FP64 atomics for the force because nothing stops a list from naming a particle twice.  None of them has a
// dPotential: MD.cpp:642-669 leaves them out of the Metropolis box move.  MODE 0 force, 1 potential.

// BOUNDARY (doBoundaryForce system.h:2334-2348, doBoundaryPotential :3137-3151; boundaryF MD.h:543-580, boundaryP
// :586-626): U = k (1/d^4 - 1/d^2) for d^2 <= 2 along one axis around a plane; c = {dim, centre, unused, k}
template <int MODE>
__global__ void __launch_bounds__(TPB) k_boundary(int nb, int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of, Geom g,
                                                  const int *__restrict__ idx, int dim, double centre, double kk, double *acc, double *partials)
{
	pdl_prologue();
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	double usum = 0;
	if (b < nb) {
		int s = slot_of[idx[b]];
		Particle p = load_particle(pos + s);
		double L = g.box[dim];
		double d = (dim == 0 ? p.x : dim == 1 ? p.y : p.z) - centre;
		d -= (d > L / 2.0) ? L : 0;
		d += (d < -L / 2.0) ? L : 0;
		double dir = (d < 0) ? -1.0 : 1.0;   // MD.h:81 #define sign(v)
		d = fabs(d);
		double d2 = d * d;
		if (MODE == 0) {
			double magnitude = 0;
			if (d2 <= 2.0) magnitude = kk * (4.0 / (d2 * d2 * d) - 2.0 / (d2 * d));
			atomicAdd(acc + dim * cap + s, magnitude * dir);
		} else {
			if (d2 <= 2.0) usum = kk * (1.0 / (d2 * d2) - 1.0 / d2);
		}
	}
	if (MODE != 0) {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
}

// ---- the kinds only MDsubstrate.cpp's switch evaluates (MDsubstrate.cpp:213-262, :476-490); forces only: neither
// dataExtraction::compute nor that driver's Metropolis sum has a term for them (doPullBeadDPotential returns 0)

// OFFSET_BOUNDARY (doOffsetBoundaryForce system.h:2351-2363; offsetBoundaryF MD.h:502-527): the BOUNDARY wall displaced by an
// offset, d = |x - centre| - offset; c = {dim, centre, offset, k}
__global__ void __launch_bounds__(TPB) k_offset_boundary(int nb, int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of, Geom g,
                                                         const int *__restrict__ idx, int dim, double centre, double offset, double kk, double *acc)
{
	pdl_prologue();
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nb) return;
	int s = slot_of[idx[b]];
	Particle p = load_particle(pos + s);
	double L = g.box[dim];
	double d = (dim == 0 ? p.x : dim == 1 ? p.y : p.z) - centre;
	d -= (d > L / 2.0) ? L : 0;
	d += (d < -L / 2.0) ? L : 0;
	double dir = (d < 0) ? -1.0 : 1.0;   // MD.h:81 #define sign(v)
	d = fabs(d) - offset;
	double d2 = d * d;
	double magnitude = 0;
	if (d2 <= 2.0) magnitude = kk * (4.0 / (d2 * d2 * d) - 2.0 / (d2 * d));
	atomicAdd(acc + dim * cap + s, magnitude * dir);
}

// RIGIDBEND (doRigidBendForce system.h:2366-2400; kmaxTorqueF MD.h:1073-1113): a torque that turns the bond first -> second
// towards a preferred direction z; c = {zx, zy, zz, k, thetaD}.  asin / pow are the CUDA library's (<= 2 ulp from glibc's).
__global__ void __launch_bounds__(TPB) k_rigidbend(int nb, int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of, Geom g,
                                                   const int *__restrict__ ij, double zx, double zy, double zz, double kk, double thetaD, double *acc)
{
	pdl_prologue();
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nb) return;
	const int s1 = slot_of[ij[2 * b]], s2 = slot_of[ij[2 * b + 1]];
	const Particle p1 = load_particle(pos + s1), p2 = load_particle(pos + s2);
	double r[3] = {p1.x - p2.x, p1.y - p2.y, p1.z - p2.z};
#pragma unroll
	for (int a = 0; a < 3; a++) {
		if (r[a] > g.box[a] / 2.0) r[a] -= g.box[a];
		if (r[a] < -g.box[a] / 2.0) r[a] += g.box[a];
	}
	double dr = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
	r[0] /= dr; r[1] /= dr; r[2] /= dr;
	const double theta = asin(-((r[0] * zx) + (r[1] * zy) + (r[2] * zz)));
	const double thetaSqr = theta * theta;
	const double hp = M_PI / 2.0;
	double magnitude;
	if (theta < thetaD)
		magnitude = kk;
	else
		magnitude = kk * (-2.0 * theta * thetaSqr + 3.0 * (hp + thetaD) * thetaSqr - 6.0 * hp * thetaD * theta + (M_PI * M_PI / 4.0) * (3.0 * thetaD - hp)) /
		            pow(thetaD - hp, 3.0);
	double cp[3] = {(r[1] * zz - r[2] * zy), -(r[0] * zz - r[2] * zx), (r[0] * zy - r[1] * zx)};
	dr = sqrt(cp[0] * cp[0] + cp[1] * cp[1] + cp[2] * cp[2]);
	cp[0] /= dr; cp[1] /= dr; cp[2] /= dr;
	const double f[3] = {(cp[1] * r[2] - cp[2] * r[1]) * magnitude, -(cp[0] * r[2] - cp[2] * r[0]) * magnitude, (cp[0] * r[1] - cp[1] * r[0]) * magnitude};
#pragma unroll
	for (int a = 0; a < 3; a++) { atomicAdd(acc + a * cap + s1, f[a]); atomicAdd(acc + a * cap + s2, -f[a]); }
}

// PULLBEAD (doPullBeadForce system.h:2449-2486; harmonicFZ MD.h:434-447): a spring between a particle and a fixed point whose
// reaction is spread over EVERY particle (the pulled one included); c = {x0, y0, z0, k}.  One thread per particle walks the
// records in order -- the record count is a handful --, so a particle's terms are added in the reference's order.
__global__ void __launch_bounds__(TPB) k_pullbead(Cnt cnt, int nTotal, int nb, int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of,
                                                  Geom g, const int *__restrict__ idx, double x0, double y0, double z0, double kk, double *acc)
{
	pdl_prologue();
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= cnt.get()) return;
	double a[3] = {0.0, 0.0, 0.0};
	const double c[3] = {x0, y0, z0};
	for (int j = 0; j < nb; j++) {
		const int sp = slot_of[idx[j]];
		const Particle p = load_particle(pos + sp);
		double d[3] = {p.x - c[0], p.y - c[1], p.z - c[2]};
#pragma unroll
		for (int q = 0; q < 3; q++) {
			if (d[q] > g.box[q] / 2.0) d[q] -= g.box[q];
			if (d[q] < -g.box[q] / 2.0) d[q] += g.box[q];
			d[q] *= kk;
			if (sp == s) a[q] -= d[q];
			a[q] += (0.0 + d[q]) / (double)(nTotal - 1);
		}
	}
	// (this kernel is the only writer of a[] while it runs: launches on one stream are ordered)
	acc[s] += a[0]; acc[cap + s] += a[1]; acc[2 * cap + s] += a[2];
}

// FLOATING_BASE (doFloatingBaseForce system.h:2402-2422, doFloatingBasePotential :2424-2446; floatingBaseForce MD.h:457-472,
// floatingBasePotential :475-494): a polynomial wall in z, constants row 6 * type.  Quirk reproduced: the force takes z
// itself, the potential z - C[0] of ROW 0.
template <int MODE>
__global__ void __launch_bounds__(TPB) k_floating_base(int nb, int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of,
                                                       const int *__restrict__ idx, const double *__restrict__ C, double *acc, double *partials)
{
	pdl_prologue();
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	double usum = 0;
	if (b < nb) {
		int s = slot_of[idx[b]];
		Particle p = load_particle(pos + s);
		const double *k = C + 6 * p.type;
		double k0 = k[0], k1 = k[1];
		if (MODE == 0) {
			double z = p.z;
			if (z <= k0) {
				atomicAdd(acc + 2 * cap + s, k[3] * z * z * z - 2 * k[3] * k0 * z * z + (2 * k[2] + k[3] * k0 * k0) * z);
			} else if (z < k1) {
				double rcz = z - k1;
				atomicAdd(acc + 2 * cap + s, k[5] * (rcz * rcz) * (2 * z * z - (4 * k1 - 7 * k0) * z + 2 * k1 * k1 + k0 * (-7 * k1 + 6 * k0)));
			}
		} else {
			double z = p.z - C[0];
			if (z <= k0) {
				double rmz = k0 - z;
				usum = k[2] * (k0 * k0 - z * z) + k[3] * rmz * rmz * rmz * ((k0 / 3.0) - (1.0 / 4.0) * rmz) + k[4];
			} else if (z < k1) {
				double rcz = k1 - z;
				usum = k[5] * rcz * rcz * rcz * (rcz * ((2.0 / 5.0) * rcz - (7.0 / 4.0) * k0) + (2.0) * k0 * k0);
			}
		}
	}
	if (MODE != 0) {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
}

// ZTORQUE (doZTorqueForce system.h:2625-2663, doZTorquePotential :2582-2623; zTorqueForce MD.h:1126-1140, zTorquePotential
// :1116-1124): aligns the bonds (l, l+1), l = k .. k+len-3, of a block of chains with z, strength eps(z) = a - b tanh(c (z - d))
// at the bond's mid height; the force switches eps off above z = d, the potential does not.  One thread per bond.
template <int MODE>
__global__ void __launch_bounds__(TPB) k_ztorque(int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of, Geom g,
                                                 int start, int nChains, int len, double c0, double c1, double c2, double c3,
                                                 double *acc, double *partials)
{
	pdl_prologue();
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	int per = len - 2;
	double usum = 0;
	if (per > 0 && t < nChains * per) {
		int l = start + (t / per) * len + (t % per);
		int s1 = slot_of[l], s2 = slot_of[l + 1];
		Particle p1 = load_particle(pos + s1);
		V3 d = diff_mi(p1, load_particle(pos + s2), g);
		double z = p1.z + d.z / 2.0;
		while (z >= g.box[2]) z -= g.box[2];
		while (z < 0) z += g.box[2];
		double mag = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
		double eps = c0 - c1 * tanh(c2 * (z - c3));
		if (MODE == 0) {
			if (z > c3) eps = 0;
			double m = -eps * (d.z / (mag * mag * mag * mag));
			double fx = d.z * m * d.x, fy = d.z * m * d.y, fz = m * (d.x * d.x + d.y * d.y);
			atomicAdd(acc + s1, fx); atomicAdd(acc + cap + s1, fy); atomicAdd(acc + 2 * cap + s1, -fz);
			atomicAdd(acc + s2, -fx); atomicAdd(acc + cap + s2, -fy); atomicAdd(acc + 2 * cap + s2, fz);
		} else {
			double cosTheta = fabs(d.z) / mag;
			usum = eps * (1.0 - cosTheta * cosTheta) / 2.0;
		}
	}
	if (MODE != 0) {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
}

// ZPOWERPOTENTIAL (doZPowerForce system.h:2692-2715, doZPowerPotential :2665-2690; zPowerForce MD.h:1148-1153,
// zPowerPotential :1142-1146): a.z += k z^n, U = -k z^(n+1) / (n+1) for the particles start .. start+count-1
template <int MODE>
__global__ void __launch_bounds__(TPB) k_zpower(int count, int cap, const Particle *__restrict__ pos, const int *__restrict__ slot_of, int start,
                                                double kk, double n, double *acc, double *partials)
{
	pdl_prologue();
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	double usum = 0;
	if (t < count) {
		int s = slot_of[start + t];
		double z = load_particle(pos + s).z;
		if (MODE == 0) atomicAdd(acc + 2 * cap + s, kk * pow(z, n));
		else usum = -kk * pow(z, n + 1) / (n + 1);
	}
	if (MODE != 0) {
		usum = block_sum(usum);
		if (threadIdx.x == 0) partials[blockIdx.x] = usum;
	}
}

// a[bead] /= 4 pi R^2 (MD.cpp:340-355, :480-494)
__global__ void k_bead_mass(int n, int cap, const int *__restrict__ beads, const int *__restrict__ slot_of, double mass, double *acc)
{
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	int s = slot_of[beads[j]];
	acc[s] /= mass; acc[cap + s] /= mass; acc[2 * cap + s] /= mass;
}

// NANOCORE: a[bead j] /= 4 pi R_j^2, R_j = C[22 j + 4] (MD.cpp:290-303, :495-508)
__global__ void k_nanocore_mass(int n, int cap, const int *__restrict__ beads, const int *__restrict__ slot_of, const double *__restrict__ C,
                                double *acc)
{
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	double R = C[22 * j + 4];
	double mass = (4.0) * M_PI * R * R;
	int s = slot_of[beads[j]];
	acc[s] /= mass; acc[cap + s] /= mass; acc[2 * cap + s] /= mass;
}

// ------------------------------------------------------------------------------------------------ FP64 pipe peak
// Roofline denominator for the pair kernel, measured on the device it runs on: 8 independent dependency chains per
// thread of either DFMA (FUSED = 1, the pipe's nominal peak, 2 flop per instruction) or DMUL + DADD (FUSED = 0, what
// this library may issue: no contraction allowed, 1 flop per instruction).
template <int FUSED>
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double a, double b, double *out)
{
	double x[8];
#pragma unroll
	for (int k = 0; k < 8; k++) x[k] = (double)(threadIdx.x + k) * 1e-3;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int k = 0; k < 8; k++) {
			if (FUSED) x[k] = __fma_rn(x[k], a, b);
			else x[k] = __dadd_rn(__dmul_rn(x[k], a), b);
		}
	}
	double s = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) s += x[k];
	if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

// ------------------------------------------------------------------------------------------------ gather / scatter to original order
__global__ void __launch_bounds__(TPB) k_export_particles(int N, int cap, const Particle *pos, const double *vel, const int *gid,
                                                          double *xyz, int *type, double *v)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= N) return;
	int g = gid[s];
	Particle p = load_particle(pos + s);
	if (xyz) { xyz[3 * g] = p.x; xyz[3 * g + 1] = p.y; xyz[3 * g + 2] = p.z; }
	if (type) type[g] = p.type;
	if (v) { v[3 * g] = vel[s]; v[3 * g + 1] = vel[cap + s]; v[3 * g + 2] = vel[2 * cap + s]; }
}

__global__ void __launch_bounds__(TPB) k_export_soa3(int N, int cap, const double *a, const int *gid, double *out)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= N) return;
	int g = gid[s];
	out[3 * g] = a[s]; out[3 * g + 1] = a[cap + s]; out[3 * g + 2] = a[2 * cap + s];
}

__global__ void __launch_bounds__(TPB) k_export_int(int N, const int *v, const int *gid, int *out)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= N) return;
	out[gid[s]] = v[s];
}

// gid_in: slab mode -- global index (+ ghost flag) of each uploaded particle; single GPU: slot s holds particle s
__global__ void __launch_bounds__(TPB) k_import_particles(int N, int cap, const double *xyz, const int *type, const double *v,
                                                          Particle *pos, double *vel, double *unw, int *gid, int *slot_of,
                                                          const int *gid_in, Geom g, int nT, int *bad, int n_global)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= N) return;
	Particle p;
	p.x = xyz[3 * s]; p.y = xyz[3 * s + 1]; p.z = xyz[3 * s + 2];
	p.type = type[s];
	p.cell = 0;
	if (bad) {   // what the reference refuses at load (system.h:452-469): the first offender by (particle, axis); 3: its type
		const double c[3] = {p.x, p.y, p.z};
		int code = INT_MAX;
		for (int d = 2; d >= 0; d--) if (!(c[d] >= 0 && c[d] <= g.box[d])) code = 4 * s + d;
		if (code == INT_MAX && (p.type < 0 || p.type >= nT)) code = 4 * s + 3;
		if (code != INT_MAX) atomicMin(bad, code);
		if (gid_in && ((gid_in[s] & GID_MASK) < 0 || (gid_in[s] & GID_MASK) >= n_global)) atomicMin(bad + 1, s);   // slab: global index out of range
	}
	store_particle(pos + s, p);
	vel[s] = v ? v[3 * s] : 0.0; vel[cap + s] = v ? v[3 * s + 1] : 0.0; vel[2 * cap + s] = v ? v[3 * s + 2] : 0.0;
	if (unw) { unw[s] = p.x; unw[cap + s] = p.y; unw[2 * cap + s] = p.z; }
	int gi = gid_in ? gid_in[s] : s;
	gid[s] = gi;
	if ((gi & GID_MASK) >= 0 && (gi & GID_MASK) < n_global) slot_of[gi & GID_MASK] = s;
}

// reference cell key and rank inside the cell's list, per original index
__global__ void __launch_bounds__(TPB) k_export_cells(int N, const Particle *pos, const int *gid,
                                                      const int *start, const int *win, Geom g, int *key, int *rank)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= N) return;
	int cx, cy, cz;
	unpack_cell(pos[s].cell, cx, cy, cz);
	int id = gid[s];
	key[id] = cx + cy * g.nc[0] + cz * g.nc[0] * g.nc[1];
	// position in the reference's head-inserted linked list of the cell (cellOpt.h:572-585): descending original index.
	// (The sorted order itself is by x slice first, see Geom::xs.)
	int c = (cx - win[WIN_ORG]) * g.xs + win[WIN_FD0] * ((cy - win[WIN_ORG + 1]) + win[WIN_DIM + 1] * (cz - win[WIN_ORG + 2]));
	int r = 0;
	for (int k = start[c]; k < start[c + g.xs]; k++) r += ((gid[k] & GID_MASK) > (id & GID_MASK));
	rank[id] = r;
}

// ------------------------------------------------------------------------------------------------ observables
// dataExtraction::compute's geometric observables reduced on the device: bond / bend means
// (dataExtraction.h:861-937), flicker = extent of the particles (:1457-1487), the kinetic-energy histogram (:1511-1520),
// the mean square displacement per molecule (:1525-1663).  Sums: one partial per block, then k_final_sum (deterministic).

// order-preserving map double -> unsigned 64-bit, so that atomicMin / atomicMax do the comparison
__device__ __forceinline__ unsigned long long obs_key(double v)
{
	unsigned long long b = (unsigned long long)__double_as_longlong(v);
	return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct ObsHist { unsigned long long *bins; long long cap; unsigned long long *overflow; int overflow_cap; int *n_overflow; };

// ext[0..2] start at the box size, ext[3..5] at 0: the reference's initial values (dataExtraction.h:1457-1460)
__global__ void k_obs_init(unsigned long long *ext, double bx, double by, double bz)
{
	ext[0] = obs_key(bx); ext[1] = obs_key(by); ext[2] = obs_key(bz);
	ext[3] = ext[4] = ext[5] = obs_key(0.0);
}

// one pass over the particles: extent (min / max per axis) and, if hist.bins, the kinetic-energy histogram
// bin = int(0.5 v^2 / partition); a bin beyond the table goes to a short overflow list the host merges
__global__ void __launch_bounds__(256) k_obs_particles(Cnt cnt, int cap, const Particle *__restrict__ pos, const double *__restrict__ vel,
                                                        const int *__restrict__ gid, unsigned long long *ext, ObsHist hist, double partition)
{
	const int N = cnt.get();
	unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
	for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < N; s += gridDim.x * blockDim.x) {
		if (gid[s] & GID_GHOST) continue;
		const Particle p = load_particle(pos + s);
		const double c[3] = {p.x, p.y, p.z};
#pragma unroll
		for (int a = 0; a < 3; a++) { unsigned long long k = obs_key(c[a]); lo[a] = min(lo[a], k); hi[a] = max(hi[a], k); }
		if (hist.bins) {
			const double vx = vel[s], vy = vel[cap + s], vz = vel[2 * cap + s];
			const long long b = (long long)(0.5 * (vx * vx + vy * vy + vz * vz) / partition);
			if (b >= 0 && b < hist.cap) atomicAdd(hist.bins + b, 1ull);
			else {
				int k = atomicAdd(hist.n_overflow, 1);
				if (k < hist.overflow_cap) hist.overflow[k] = (unsigned long long)b;
			}
		}
	}
#pragma unroll
	for (int a = 0; a < 3; a++) {
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], d));
			hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], d));
		}
		if ((threadIdx.x & 31) == 0) {
			if (lo[a] != ~0ull) atomicMin(ext + a, lo[a]);
			if (hi[a] != 0ull) atomicMax(ext + 3 + a, hi[a]);
		}
	}
}

// the reference's image rule for its bond statistics (dataExtraction.h:876-884): both comparisons inclusive
__device__ __forceinline__ double obs_image(double d, double L)
{
	if (d >= L / 2.0) d -= L;
	if (d <= -L / 2.0) d += L;
	return d;
}

// BOND records: sum of the bond lengths (dataExtraction.h:861-893)
__global__ void __launch_bounds__(TPB) k_obs_bond(int n, const int *__restrict__ ij, const Particle *__restrict__ pos,
                                                  const int *__restrict__ slot_of, Geom g, double *partials)
{
	const int l = blockIdx.x * blockDim.x + threadIdx.x;
	double r = 0.0;
	if (l < n) {
		const Particle a = load_particle(pos + slot_of[ij[2 * l]]), b = load_particle(pos + slot_of[ij[2 * l + 1]]);
		const double dx = obs_image(a.x - b.x, g.box[0]), dy = obs_image(a.y - b.y, g.box[1]), dz = obs_image(a.z - b.z, g.box[2]);
		r = sqrt(dx * dx + dy * dy + dz * dz);
	}
	r = block_sum(r);
	if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

// BEND records: sums of cos(theta) and of the two arm lengths (dataExtraction.h:895-935); partials[q * gridDim.x + block]
__global__ void __launch_bounds__(TPB) k_obs_bend(int n, const int *__restrict__ ijk, const Particle *__restrict__ pos,
                                                  const int *__restrict__ slot_of, Geom g, double *partials)
{
	const int l = blockIdx.x * blockDim.x + threadIdx.x;
	double ct = 0.0, ra = 0.0, rb = 0.0;
	if (l < n) {
		const Particle a = load_particle(pos + slot_of[ijk[3 * l]]), b = load_particle(pos + slot_of[ijk[3 * l + 1]]),
		               c = load_particle(pos + slot_of[ijk[3 * l + 2]]);
		const double ax = obs_image(a.x - b.x, g.box[0]), ay = obs_image(a.y - b.y, g.box[1]), az = obs_image(a.z - b.z, g.box[2]);
		const double bx = obs_image(b.x - c.x, g.box[0]), by = obs_image(b.y - c.y, g.box[1]), bz = obs_image(b.z - c.z, g.box[2]);
		ra = sqrt(ax * ax + ay * ay + az * az);
		rb = sqrt(bx * bx + by * by + bz * bz);
		ct = (ax * bx + ay * by + az * bz) / (ra * rb);
	}
	ct = block_sum(ct);
	if (threadIdx.x == 0) partials[blockIdx.x] = ct;
	ra = block_sum(ra);
	if (threadIdx.x == 0) partials[gridDim.x + blockIdx.x] = ra;
	rb = block_sum(rb);
	if (threadIdx.x == 0) partials[2 * gridDim.x + blockIdx.x] = rb;
}

// aPStart (MD.cpp:96-105): the unwrapped positions at the start of the diffusion measurement, by original index, SoA [3][N]
__global__ void __launch_bounds__(TPB) k_obs_msd_start(int N, int cap, const double *__restrict__ unw, const int *__restrict__ gid, double *start)
{
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= N) return;
	const int id = gid[s] & GID_MASK;
	start[id] = unw[s]; start[N + id] = unw[cap + s]; start[2 * N + id] = unw[2 * cap + s];
}

// sum of |unwrapped - start|^2 over the entries of one molecule: idx != null: the list idx[0 .. n) of original indices (an
// index listed twice counts twice, as in the reference's loop over the records); else the range first .. first + n
__global__ void __launch_bounds__(TPB) k_obs_msd(int n, const int *__restrict__ idx, int first, int N, int cap, const double *__restrict__ unw,
                                                 const int *__restrict__ slot_of, const double *__restrict__ start, double *partials)
{
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	double r = 0.0;
	if (e < n) {
		const int id = idx ? idx[e] : first + e;
		const int s = slot_of[id];
		const double d0 = unw[s] - start[id], d1 = unw[cap + s] - start[N + id], d2 = unw[2 * cap + s] - start[2 * N + id];
		r = d0 * d0 + d1 * d1 + d2 * d2;
	}
	r = block_sum(r);
	if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

} // namespace smd
