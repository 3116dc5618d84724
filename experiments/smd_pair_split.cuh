// The pair engine as TWO kernels (included by smd_core.cu after smd_kernels.cuh) -- an EXPERIMENT that lost, kept
// selectable (SMD_PAIR_SPLIT=1) and parity-tested because the next round's work on the pair kernel starts from it.
//
// k_pair_force2 (smd_kernels.cuh) does the FP32 prefilter and the FP64 evaluation in one kernel with the candidate
// lists in shared memory.  Its two phases want opposite things: the prefilter is a stream of divergent 16-byte gathers
// and needs many warps but few registers; the evaluation needs ~100 registers for its interleaved FP64 chains.  In one
// kernel both run at the occupancy of the hungrier one (126 registers, 16 warps per SM; clock64 instrumentation,
// SMD_EXP_TIMING: 10 k cycles of set-up, 42 k of prefilter and 26 k of evaluation per block).  Split:
//
//   k_pair_lists   one thread per particle: candidate ranges, FP32 prefilter, 16-bit list entries (range id, offset)
//                  written to the particle's row of a global list buffer, 80 registers, 6 blocks per SM.  Pairs through
//                  a periodic image and list overflow are evaluated here in FP64 by the out-of-line general routines
//                  (rare) and handed over as partial sums.
//   k_pair_drain   per block of 128 particles: lists re-dealt by length (as before), FP64 evaluation with the entries
//                  read back through L1, Langevin term / reductions in the epilogue; 95 registers, 5 blocks per SM.
//
// Same pairs, same per-pair arithmetic, same summation order per particle as k_pair_force2: results are bit-identical
// (tests/test_gpu_parity.py).  Measured on C2 (B200): 237 us for the two kernels against 177 us for k_pair_force2.
// ncu on k_pair_force2 explains it: the L1 data pipe is the busiest unit (62 % of its wavefront slots: 6.3 wavefronts
// per divergent 16-byte gather request), and the scattered 2-byte list stores to global memory add a second stream of
// one-sector wavefronts to that pipe, while the one-kernel version keeps them in shared memory and lets blocks in
// different phases overlap on different units.  See DESIGN.md section 3.2.
#pragma once
#include "smd_kernels.cuh"

namespace smd {

constexpr int NL_CAP = 160;               // list entries per particle (16 bit each); multiple of 8
constexpr int NL_FLAG_PART = 1 << 30;     // nl.cnt flag: this particle has partial sums in nl.part

struct NeighLists {
	unsigned short *ent;   // [cap][NL_CAP]
	int *rng;              // [PAIR_NSEG][cap] first slot of each candidate range
	int *cnt;              // [cap] list length (| NL_FLAG_PART)
	double *part;          // [3][cap] sums that bypass the list
};

struct ListSmem {
	int perm[PAIR_TPB];
	int wcnt[PAIR_TPB / 32];
	int seg_b[PAIR_NSEG][PAIR_TPB];
	unsigned short seg_n[PAIR_NSEG][PAIR_TPB];
};

#ifndef SMD_LISTS_BLOCKS
#define SMD_LISTS_BLOCKS 6
#endif
#ifndef SMD_DRAIN_BLOCKS
#define SMD_DRAIN_BLOCKS 5
#endif

template <int EMODE>
__global__ void __launch_bounds__(PAIR_TPB, SMD_LISTS_BLOCKS) k_pair_lists(Cnt cnt, int cap, const Particle *__restrict__ pos, const float4 *__restrict__ pos32,
                                                            const int *__restrict__ start, const int *__restrict__ win, Geom g, int nT,
                                                            const double *__restrict__ tab, PairGeo pg, const int *__restrict__ gid,
                                                            EnergyArgs en, NeighLists nl)
{
	const int N = cnt.get();
	if ((int)(blockIdx.x * PAIR_TPB) >= N) return;
	__shared__ ListSmem sm;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

	// ---- deal the block's particles to threads by class (long-range types first): warps of similar work
	{
		const int i0 = blockIdx.x * PAIR_TPB + tid;
		const bool heavy = (i0 < N) && pos32[i0].w >= pg.thr32;
		const unsigned bal = __ballot_sync(0xffffffffu, heavy);
		if (lane == 0) sm.wcnt[wid] = __popc(bal);
		__syncthreads();
		int before = 0, total = 0;
#pragma unroll
		for (int k = 0; k < PAIR_TPB / 32; k++) { int c = sm.wcnt[k]; total += c; if (k < wid) before += c; }
		const int below = __popc(bal & ((1u << lane) - 1u));
		const int rank = heavy ? before + below : total + (tid - before - below);
		sm.perm[rank] = i0;
		__syncthreads();
	}
	const int i = sm.perm[tid];
	if (i >= N) return;
	const bool live = !(g.slab && (gid[i] & GID_GHOST));   // slab mode: ghosts are only neighbours
	if (!live) { nl.cnt[i] = 0; return; }

	const float4 p32 = pos32[i];
	int cx, cy, cz;
	unpack_cell(pos[i].cell, cx, cy, cz);
	const int w0 = win[WIN_ORG], w1 = win[WIN_ORG + 1], w2 = win[WIN_ORG + 2];
	const int d0 = win[WIN_DIM], d1 = win[WIN_DIM + 1], d2 = win[WIN_DIM + 2];
	const float ai = p32.w;
	const float ext = EMODE != 0 ? en.extra32 : 0.f;
	double ex = 0, ey = 0, ez = 0;               // sums that bypass the list
	bool have_part = false;
	// one candidate evaluated on the spot in FP64 (periodic images, overflow): the general out-of-line routines
	auto direct = [&](int j, bool unshifted_once) {
		const Particle pi = load_particle(pos + i), pj = load_particle(pos + j);
		if (EMODE == 0) {
			D3 f = pair_force_term(i, pi, j, pj, g, nT, tab, 6 * nT * nT);
			ex += f.x; ey += f.y; ez += f.z;
		} else {
			ex += pair_energy_term<(EMODE ? EMODE : 1)>(i, pi, j, pj, g, nT, tab, en.sx, en.sy, en.sz, unshifted_once);
		}
		have_part = true;
	};
	auto near32 = [&](float qx, float qy, float qz, const float4 &c) {
		float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
		return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)) < fminf(ai, c.w) + ext;
	};

	// ---- the particle's candidate ranges: one per (y,z) row of the stencil, pruned by geometry (see k_pair_force2)
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
	const float amax = fminf(ai, pg.thr32) + ext;
	int nseg = 0;
	bool shifted_rows = false;
#pragma unroll 1
	for (int r = 0; r < 9; r++) {
		int oz = r / 3 - 1, oy = r - 3 * (r / 3) - 1;
		int nz = cz + oz, ny = cy + oy;
		bool wrapyz = nz < 0 || nz >= g.nc[2] || ny < 0 || ny >= g.nc[1];
		int lz = nz - w2, ly = ny - w1;
		float gy = oy == 0 ? 0.f : (oy > 0 ? fp[1] : fm[1]), gz = oz == 0 ? 0.f : (oz > 0 ? fp[2] : fm[2]);
		float gyz = gy * gy + gz * gz;
		bool row_ok = gyz < amax;
		if (row_ok && (wrapyz || cx == 0 || cx == g.nc[0] - 1)) shifted_rows = true;
		if (EMODE != 0 && (oz < 0 || (oz == 0 && oy < 0))) row_ok = false;   // backward rows: the other particle counts the pair
		row_ok = row_ok && !wrapyz && lz >= 0 && lz < d2 && ly >= 0 && ly < d1;
		bool keep_lo = gyz + fm[0] * fm[0] < amax, keep_hi = gyz + fp[0] * fp[0] < amax;
		int xlo, xhi;
		if (g.slab) {
			xlo = win_x(max(cx - (keep_lo ? 1 : 0), 0), w0, g.nc[0]); xhi = win_x(min(cx + (keep_hi ? 1 : 0), g.nc[0] - 1), w0, g.nc[0]);
		} else {
			xlo = max(max(cx - (keep_lo ? 1 : 0), 0) - w0, 0); xhi = min(min(cx + (keep_hi ? 1 : 0), g.nc[0] - 1) - w0, d0 - 1);
		}
		int jb = 0, je = 0;
		if (row_ok && xlo <= xhi) {
			int rowbase = d0 * (ly + d1 * lz);
			jb = start[rowbase + xlo]; je = start[rowbase + xhi + 1];
			if (EMODE != 0 && oz == 0 && oy == 0) jb = min(max(jb, i + 1), je);   // own row: only the slots behind mine
		}
		sm.seg_b[r][tid] = jb;
		nl.rng[r * cap + i] = jb;
		const int lim = (1 << PAIR_SEGBITS) - 4;
		sm.seg_n[r][tid] = (unsigned short)min(je - jb, lim);
		// a range longer than the 12-bit offset field (> 1300 particles per cell): take the excess one by one
		for (int j = jb + lim; j < je; j++)
			if (near32(p32.x, p32.y, p32.z, pos32[j]) && j != i) direct(j, true);
		nseg = (je > jb) ? r + 1 : nseg;
	}

	// ---- phase 1: FP32 prefilter along the ranges; four candidates per group, the next group in flight
	unsigned short *const row = nl.ent + (size_t)i * NL_CAP;
	unsigned short *wp = row;
	unsigned short *const wlim = row + (NL_CAP - 8);     // checked once per two groups of four
	int sg = 0, q = 0;
	bool overflow = false;
	for (; sg < nseg && !overflow; sg++) {
		const int n = sm.seg_n[sg][tid];
		if (n == 0) continue;
		const float4 *cp = pos32 + sm.seg_b[sg][tid];
		const unsigned tag = (unsigned)sg << PAIR_SEGBITS;
		float4 ga[4], gb[4];                         // ping-pong buffers: one group under test, the next in flight
#pragma unroll
		for (int k = 0; k < 4; k++) ga[k] = cp[k];   // pos32 is padded: the overhang is masked below
		q = 0;
		auto group = [&](const float4 (&c)[4], float4 (&nx)[4]) {
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
				float dx = p32.x - c[k].x, dy = p32.y - c[k].y, dz = p32.z - c[k].z;
				float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
				const float thr = EMODE != 0 ? fminf(ai, c[k].w) + ext : fminf(ai, c[k].w);
				if (r2 < thr && k < rem) *wp++ = (unsigned short)(e0 + k);
			}
			return more;
		};
		while (true) {
			if (!group(ga, gb)) break;
			if (!group(gb, ga)) break;
			if (wp > wlim) { overflow = true; break; }
		}
		if (!overflow && wp > wlim) { overflow = true; q = n; }
	}
	if (overflow) {
		// the list is full (never seen with NL_CAP = 160): the candidates not yet examined -- the rest of range sg - 1
		// from offset q on, and all later ranges -- are evaluated on the spot
		for (int s2 = sg - 1; s2 < nseg; s2++) {
			const int n = sm.seg_n[s2][tid], jb = sm.seg_b[s2][tid];
			for (int j = (s2 == sg - 1 ? q : 0); j < n; j++)
				if (near32(p32.x, p32.y, p32.z, pos32[jb + j]) && jb + j != i) direct(jb + j, true);
		}
	}

	// ---- rows and end cells seen through a periodic image (particles in the outermost cell layers only)
	if (shifted_rows) {
#pragma unroll 1
		for (int s = 0; s < 27; s++) {
			int r = s / 3, sub = s - 3 * r;
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
			int rowbase = d0 * (ly + d1 * lz);
			int jb = start[rowbase + xlo], je = start[rowbase + xhi + 1];
			for (int j = jb; j < je; j++)
				if (near32(p32.x - sx, p32.y - sy, p32.z - sz, pos32[j])) direct(j, false);
		}
	}

	int c = (int)(wp - row);
	if (have_part) {
		nl.part[i] = ex; nl.part[cap + i] = ey; nl.part[2 * cap + i] = ez;
		c |= NL_FLAG_PART;
	}
	nl.cnt[i] = c;
}

struct DrainSmem {
	int cnt[PAIR_TPB];                  // list length (flag stripped) of the block's particles
	int flag[PAIR_TPB];
	int order[PAIR_TPB];                // thread -> particle (offset in the block) whose list it drains
	int seg_b[PAIR_NSEG][PAIR_TPB];     // candidate range starts of each particle
	int hist[NL_CAP + 2];
	double usum[PAIR_TPB];              // energy modes: per-particle sums, reduced in particle order
};

template <int EMODE, bool LANGEVIN, bool SYMM>
__global__ void __launch_bounds__(PAIR_TPB, SMD_DRAIN_BLOCKS) k_pair_drain(Cnt cnt, int cap, const Particle *__restrict__ pos, Geom g, int nT,
                                                            const double *__restrict__ tab, const double *__restrict__ ptab,
                                                            double *__restrict__ acc, LangevinArgs lg, const int *__restrict__ gid,
                                                            EnergyArgs en, NeighLists nl)
{
	const int N = cnt.get();
	if ((int)(blockIdx.x * PAIR_TPB) >= N) {
		if (EMODE != 0 && threadIdx.x == 0) en.partials[blockIdx.x] = 0.0;
		return;
	}
	extern __shared__ __align__(16) unsigned char s_raw[];
	double *s_ptab = reinterpret_cast<double *>(s_raw);
	__shared__ DrainSmem sm;
	const int nptab = PTAB_STRIDE * nT * nT;
	const int tid = threadIdx.x;
	const int i0 = blockIdx.x * PAIR_TPB;
	for (int k = tid; k < nptab; k += PAIR_TPB) s_ptab[k] = ptab[k];
	for (int k = tid; k < NL_CAP + 2; k += PAIR_TPB) sm.hist[k] = 0;
	int lcnt = 0;
	{
		const int i = i0 + tid;
		int c = (i < N) ? nl.cnt[i] : 0;
		lcnt = c & 0xffff;
		sm.cnt[tid] = lcnt;
		sm.flag[tid] = c & NL_FLAG_PART;
		if (lcnt > 0) {
#pragma unroll
			for (int r = 0; r < PAIR_NSEG; r++) sm.seg_b[r][tid] = nl.rng[r * cap + i];
		}
	}
	__syncthreads();
	// ---- hand the lists out sorted by length, longest first (counting sort), so that the lanes of a warp drain lists
	// of nearly equal length
	atomicAdd(&sm.hist[lcnt], 1);
	__syncthreads();
	if (tid < 32) {
		constexpr int PER = (NL_CAP + 1 + 31) / 32;
		int h[PER], sum = 0;
#pragma unroll
		for (int k = 0; k < PER; k++) { int c = NL_CAP - (tid * PER + k); h[k] = c >= 0 ? sm.hist[c] : 0; sum += h[k]; }
		int inc = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, d); if (tid >= d) inc += v; }
		int run = inc - sum;
#pragma unroll
		for (int k = 0; k < PER; k++) { int c = NL_CAP - (tid * PER + k); if (c >= 0) sm.hist[c] = run; run += h[k]; }
	}
	__syncthreads();
	sm.order[atomicAdd(&sm.hist[lcnt], 1)] = tid;
	__syncthreads();

	const int o = sm.order[tid];
	const int io = i0 + o;
	const bool act = io < N && !(g.slab && (gid[io] & GID_GHOST));
	if (EMODE == 0 && !act) return;
	double ax = 0, ay = 0, az = 0;
	if (act) {
		const Particle po = load_particle(pos + io);
		if (sm.flag[o]) { ax = nl.part[io]; ay = nl.part[cap + io]; az = nl.part[2 * cap + io]; }
		const double rc2 = g.rc2;
		const char *rowi = reinterpret_cast<const char *>(s_ptab + PTAB_STRIDE * po.type * nT);
		const int *segcol = &sm.seg_b[0][o];
		const unsigned short *rp = nl.ent + (size_t)io * NL_CAP;
		const unsigned short *const wend = rp + sm.cnt[o];
		auto fetch = [&](const unsigned short *a, int &j) {       // entry -> neighbour slot and record
			const unsigned e = __ldg(a);
			j = segcol[(e >> PAIR_SEGBITS) * PAIR_TPB] + (int)(e & ((1u << PAIR_SEGBITS) - 1u));
			return load_particle(pos + j);
		};
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
			return in && !ok;
		};
		auto general = [&](int j, const Particle &pj) {
			D3 f = pair_force_term(io, po, j, pj, g, nT, tab, 6 * nT * nT);
			ax += f.x; ay += f.y; az += f.z;
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
		if (EMODE != 0) {
			while (rp < wend) {
				int j0;
				Particle p0 = fetch(rp, j0);
				rp++;
				if (rp < wend) {
					int j1;
					Particle p1 = fetch(rp, j1);
					rp++;
					efast(j0, p0);
					efast(j1, p1);
				} else {
					efast(j0, p0);
				}
			}
		} else {
			if (rp + 1 < wend) {
				int j0, j1;
				Particle p0 = fetch(rp, j0), p1 = fetch(rp + 1, j1);
				rp += 2;
				while (true) {
					int n0 = j0, n1 = j1;
					Particle q0 = p0, q1 = p1;
					const bool more = rp + 1 < wend;
					if (more) { q0 = fetch(rp, n0); q1 = fetch(rp + 1, n1); }   // next two pairs: in flight during the math below
					bool s0 = fast(j0, p0);
					bool s1 = fast(j1, p1);
					if (s0 || s1) {
						if (s0) general(j0, p0);
						if (s1) general(j1, p1);
					}
					if (!more) break;
					rp += 2;
					j0 = n0; j1 = n1; p0 = q0; p1 = q1;
				}
			}
			if (rp < wend) {
				int j0;
				Particle p0 = fetch(rp, j0);
				if (fast(j0, p0)) general(j0, p0);
			}
		}
	}
	if (EMODE != 0) {   // one partial sum per block, reduced deterministically by k_final_sum
		// summed in particle order, not in the (arrival-dependent) order the lists were handed out in: the energy is
		// reproducible to the last bit
		__syncthreads();
		sm.usum[o] = act ? ax : 0.0;
		__syncthreads();
		double tot = block_sum(sm.usum[tid]);
		if (tid == 0) en.partials[blockIdx.x] = tot;
		return;
	}
	if (LANGEVIN) {
		int id = lg.gid[io] & GID_MASK;
		double u[3];
		if (lg.ext_noise) {
			u[0] = lg.ext_noise[3 * id]; u[1] = lg.ext_noise[3 * id + 1]; u[2] = lg.ext_noise[3 * id + 2];
		} else {
			philox_uniform3(lg.seed, lg.step, (uint32_t)id, u);
		}
		double lx = -lg.gamma * lg.vel[io] + lg.sigma * (2.0 * u[0] - 1.0);
		double ly = -lg.gamma * lg.vel[cap + io] + lg.sigma * (2.0 * u[1] - 1.0);
		double lz = -lg.gamma * lg.vel[2 * cap + io] + lg.sigma * (2.0 * u[2] - 1.0);
		acc[io] = lx + ax; acc[cap + io] = ly + ay; acc[2 * cap + io] = lz + az;
	} else {
		acc[io] += ax; acc[cap + io] += ay; acc[2 * cap + io] += az;
	}
}

} // namespace smd
