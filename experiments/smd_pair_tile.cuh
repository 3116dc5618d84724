// The pair engine of round 2: k_pair_tile.  CellOpt::computeForce (cellOpt.h:712-926) + Force<T> (MD.h:795-848), and
// the dPotential of a box move summed along with it (cellOpt.h:1043-1180).  Included by smd_core.cu after smd_kernels.cuh.
//
// Same pairs, same per-pair FP64 arithmetic and the same summation order as k_pair_force2 (every particle gathers its own
// force from the full 27-cell stencil: no atomics, deterministic, bit-identical to that kernel), but the prefilter
// ("phase 1") is no longer one thread walking its own candidates.  k_pair_force2 spent 40 k of its 73 k cycles per block
// there, ~20 instructions per candidate at 17.7 of 32 lanes.  Here a WARP takes a group = the particles of one reference
// cell among its 32 slots, and
//   * the group's candidates -- the nine (y,z) rows of the stencil, three cells each, concatenated -- sit one per LANE:
//     a 4-byte record {x,y,z: 8-bit coordinates inside the particle's own cell (TILE_M steps per cell edge); class, cx & 3},
//     moved into the group's frame (origin = corner of cell (cx-1, cy-1, cz-1)) by one integer add;
//   * the group's particles are looped over with their frame coordinates broadcast from shared memory; one test is
//     r^2 = |a|^2 + |b|^2 - 2 a.b with a.b ONE dp4a, then an integer multiply-add and a compare: 3 instructions for 32
//     candidate tests, plus the ballot that turns the 32 verdicts into one mask word kept by the lane of that particle;
//   * the class of a particle (long range = rc, short = purely repulsive types that see nothing beyond rm, none) rides in
//     the fourth byte of the dot product: a4 * b4 = wA * wB only when both are long range, which moves the threshold by
//     the difference of the two cutoffs -- no second compare.
// The test is conservative: coordinates are floored to steps of cs / TILE_M, a difference of two is off by less than one
// step per axis, so |dq| <= r / s_min + sqrt(3); the cutoffs are ((R / s_min) + 1.75)^2 and phase 2 repeats the exact
// FP64 test of the reference on every survivor.
// Phase 2 is k_pair_force2's: after a block barrier the 128 bit-mask lists are handed out sorted by length, each thread
// walks the set bits of one list (candidate index -> slot through the group's nine row offsets), two pairs interleaved,
// the next two records in flight.
#pragma once

namespace smd {

#ifndef SMD_TILE_BLOCKS
#define SMD_TILE_BLOCKS 4
#endif
constexpr int TILE_KW = 24;          // mask words per particle: groups with up to 768 candidates (typical: 340)
constexpr int TILE_BINS = 128;       // list lengths are sorted into this many bins
constexpr int TILE_HUGE = 1 << 29;
constexpr unsigned char TILE_OVERFLOW = 0xff;

struct TileGeo { int cut_long, cut_short, wA, wB; };   // cutoffs in steps^2, class weights (2 wA wB <= cut_long - cut_short)

struct TileSmem {
	int cnt[PAIR_TPB];                 // phase-1 survivors per particle
	int order[PAIR_TPB];               // phase-2 thread -> particle whose list it drains
	double part[3][PAIR_TPB];          // sums that bypass the lists (periodic images)
	int g_dlt[PAIR_NSEG][PAIR_TPB];    // per group (at the block-local index of its first particle), per row: slot - candidate index
	int g_end[PAIR_NSEG][PAIR_TPB];    // candidate index where the row ends
	int2 iop[PAIR_TPB / 32][32];       // per warp: {frame coordinates + class weight, threshold} of the group's particles
	int hist[TILE_BINS + 2];
	unsigned char hd[PAIR_TPB];        // block-local index of the particle's group head
	unsigned char kw[PAIR_TPB];        // mask words of the particle's group; TILE_OVERFLOW: too many candidates, brute force
	unsigned mask[TILE_KW][PAIR_TPB];  // bit c of word c / 32: candidate c of the group passed the test for this particle
};

template <int EMODE, bool LANGEVIN, bool SYMM, bool P1ONLY = false>
__global__ void __launch_bounds__(PAIR_TPB, P1ONLY ? 12 : SMD_TILE_BLOCKS) k_pair_tile(Cnt cnt, int cap, const Particle *__restrict__ pos,
                                                                          const float4 *__restrict__ pos32, const unsigned *__restrict__ pos8,
                                                                          const int *__restrict__ start, const int *__restrict__ win, Geom g,
                                                                          int nT, const double *__restrict__ tab, const double *__restrict__ ptab,
                                                                          PairGeo pg, TileGeo tg, double *__restrict__ acc, LangevinArgs lg,
                                                                          const int *__restrict__ gid, EnergyArgs en)
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
	static_assert(EMODE == 0 || EMODE == 3, "forces, or forces + the dPotential of a box move");
	static_assert(EMODE != 3 || SYMM, "forces + dPotential in one pass: symmetric tables only");
	constexpr bool DU = (EMODE == 3);
	constexpr unsigned FULL = 0xffffffffu;
	const int N = cnt.get();
	const int bid = (int)blockIdx.x;
	if (bid * PAIR_TPB >= N) {
		if (DU && threadIdx.x == 0) en.partials[blockIdx.x] = 0.0;
		return;
	}
	if (pg.done) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see k_pair_force2
	extern __shared__ __align__(16) unsigned char s_raw[];
	TileSmem &sm = *reinterpret_cast<TileSmem *>(s_raw);
	double *s_ptab = reinterpret_cast<double *>(s_raw + ((sizeof(TileSmem) + 15) & ~size_t(15)));
	const int nptab = PTAB_STRIDE * nT * nT;
	double *s_utab = s_ptab + nptab;   // DU only
	double *s_dup = s_utab + nptab;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	for (int k = tid; k < nptab; k += PAIR_TPB) s_ptab[k] = ptab[k];
	if (DU) for (int k = tid; k < nptab; k += PAIR_TPB) s_utab[k] = en.utab[k];
	for (int k = tid; k < TILE_BINS + 2; k += PAIR_TPB) sm.hist[k] = 0;
	const int w0 = win[WIN_ORG], w1 = win[WIN_ORG + 1], w2 = win[WIN_ORG + 2];
	const int d0 = win[WIN_DIM], d1 = win[WIN_DIM + 1], d2 = win[WIN_DIM + 2];
	const int fd0 = win[WIN_FD0], xs = g.xs;

	const int base = bid * PAIR_TPB;
	const int i = base + tid;
	const bool valid = i < N;
	const bool live = valid && !(g.slab && (gid[i] & GID_GHOST));   // slab mode: ghosts are only neighbours
	const double rc2 = g.rc2;

	// =============================================================== phase 1: one warp, its 32 slots, group by group
	{
		const unsigned cellw = valid ? pos[i].cell : 0u;
		const unsigned prevc = __shfl_up_sync(FULL, cellw, 1);
		const bool head = valid && (lane == 0 || cellw != prevc);
		unsigned heads = __ballot_sync(FULL, head);
		const int nval = __popc(__ballot_sync(FULL, valid));
		const unsigned livemask = __ballot_sync(FULL, live);
		const unsigned w8 = valid ? pos8[i] : 0u;
		const int wbase = wid * 32;
		if (!valid) { sm.cnt[tid] = 0; sm.kw[tid] = 0; sm.hd[tid] = (unsigned char)tid; }
		while (heads) {
			const int g0 = __ffs(heads) - 1;
			heads &= heads - 1;
			const int g1 = heads ? __ffs(heads) - 1 : nval;
			const int ni = g1 - g0;
			const int hb = wbase + g0;
			const bool mine = lane >= g0 && lane < g1;
			if (mine) sm.hd[tid] = (unsigned char)hb;
			const unsigned gm = (ni == 32 ? FULL : ((1u << ni) - 1u)) << g0;
			if (!(livemask & gm)) {   // a cell of ghosts: nobody gathers
				if (mine) { sm.cnt[tid] = 0; sm.kw[tid] = 0; }
				continue;
			}
			int cx, cy, cz;
			unpack_cell(__shfl_sync(FULL, cellw, g0), cx, cy, cz);
			// the nine rows of the stencil: slots [jb, je) of the cells cx-1 .. cx+1 that need no periodic image, clamped to
			// the window (cells outside it are empty); rows and end cells seen through an image: the per-particle loop below
			int jb = 0, je = 0;
			if (lane < PAIR_NSEG) {
				const int oz = lane / 3 - 1, oy = lane - 3 * (lane / 3) - 1;
				const int nz = cz + oz, ny = cy + oy;
				const bool wrapyz = nz < 0 || nz >= g.nc[2] || ny < 0 || ny >= g.nc[1];
				const int lz = nz - w2, ly = ny - w1;
				const bool row_ok = !wrapyz && lz >= 0 && lz < d2 && ly >= 0 && ly < d1;
				int xlo, xhi;
				if (g.slab) {
					xlo = win_x(max(cx - 1, 0), w0, g.nc[0]); xhi = win_x(min(cx + 1, g.nc[0] - 1), w0, g.nc[0]);
				} else {
					xlo = max(max(cx - 1, 0) - w0, 0); xhi = min(min(cx + 1, g.nc[0] - 1) - w0, d0 - 1);
				}
				if (row_ok && xlo <= xhi) {
					const int rowbase = fd0 * (ly + d1 * lz);
					jb = start[rowbase + xlo * xs]; je = start[rowbase + (xhi + 1) * xs];
				}
			}
			const int n = je - jb;
			int inc = n;
#pragma unroll
			for (int d = 1; d < 16; d <<= 1) { const int v = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += v; }
			const int C = __shfl_sync(FULL, inc, PAIR_NSEG - 1);
			if (lane < PAIR_NSEG) { sm.g_dlt[lane][hb] = jb - (inc - n); sm.g_end[lane][hb] = inc; }
			const int K = (C + 31) >> 5;
			if (K > TILE_KW) {   // more candidates than the masks hold (> 768 around one cell): exact brute force in phase 2
				if (mine) { sm.cnt[tid] = TILE_BINS - 1; sm.kw[tid] = TILE_OVERFLOW; }
				__syncwarp();
				continue;
			}
			if (mine) {   // own operands: frame coordinates (the own cell is the centre cell of the frame), class weight, threshold
				const unsigned cls = (w8 >> 24) & 3u;
				const unsigned a3 = (w8 & 0xffffffu) + (unsigned)(TILE_M | (TILE_M << 8) | (TILE_M << 16));
				const int na = (int)__dp4a(a3, a3, 0u);
				int thr = cls == 0u ? na - tg.cut_long + 2 * tg.wA * tg.wB : na - tg.cut_short;
				if (cls == 2u || !live) thr = INT_MAX;
				sm.iop[wid][lane - g0] = make_int2((int)(a3 | (cls == 0u ? (unsigned)tg.wA << 24 : 0u)), thr);
			}
			__syncwarp();
			int r = 0, lim = sm.g_end[0][hb], dl = sm.g_dlt[0][hb];
			int gcnt = 0;
			const int cxm1 = cx - 1;
			for (int k = 0; k < K; k++) {
				const int c = (k << 5) + lane;
				const bool vc = c < C;
				while (vc && c >= lim) { r++; lim = sm.g_end[r][hb]; dl = sm.g_dlt[r][hb]; }
				const unsigned wj = vc ? pos8[dl + c] : 0u;
				const int oz = (r * 11) >> 5, oy = r - 3 * oz;   // r / 3, r % 3 for r < 9
				const unsigned xo = ((wj >> 26) - (unsigned)cxm1) & 3u;
				const unsigned off = (xo * TILE_M) | ((unsigned)(oy * TILE_M) << 8) | ((unsigned)(oz * TILE_M) << 16);
				const unsigned b3 = (wj & 0xffffffu) + off;
				const unsigned clsj = (wj >> 24) & 3u;
				int nb = (int)__dp4a(b3, b3, 0u);
				if (!vc || clsj == 2u) nb = TILE_HUGE;
				const unsigned b = b3 | (clsj == 0u ? (unsigned)tg.wB << 24 : 0u);
				unsigned *mrow = &sm.mask[k][hb];
				for (int ii = 0; ii < ni; ii++) {
					const int2 op = sm.iop[wid][ii];
					const int t2 = (int)__dp4a((unsigned)op.x, b, 0u);
					const unsigned m = __ballot_sync(FULL, 2 * t2 - nb > op.y);
					if (lane == 0) mrow[ii] = m;   // (a loop-invariant predicate: one predicated store per test row)
				}
			}
			__syncwarp();
			if (lane < ni) {
				for (int k = 0; k < K; k++) gcnt += __popc(sm.mask[k][hb + lane]);
				sm.cnt[hb + lane] = gcnt; sm.kw[hb + lane] = (unsigned char)K;
			}
			__syncwarp();
		}
	}

	if (P1ONLY) {   // experiment: the prefilter alone, at the occupancy its own register count allows
		__syncthreads();
		if (valid) acc[i] = (double)sm.cnt[tid];
		return;
	}
	// =============================================================== pairs seen through a periodic image (own particle)
	double du = 0.0;   // DU: this thread's share of the dPotential
	double ex = 0, ey = 0, ez = 0;
	{
		Particle pi;
		pi.x = pi.y = pi.z = 0; pi.type = 0; pi.cell = 0;
		int cx = 1, cy = 1, cz = 1;
		if (live) { pi.cell = pos[i].cell; unpack_cell(pi.cell, cx, cy, cz); }
		const bool edge = live && (cx == 0 || cx == g.nc[0] - 1 || cy == 0 || cy == g.nc[1] - 1 || cz == 0 || cz == g.nc[2] - 1);
		if (edge) {
			pi = load_particle(pos + i);
			const float4 p32 = pos32[i];
			const float ai = p32.w;
			float fm[3], fp[3];
			{
				const float c[3] = {p32.x, p32.y, p32.z};
				const int ci[3] = {cx, cy, cz};
#pragma unroll
				for (int a = 0; a < 3; a++) {
					fm[a] = fmaxf(c[a] - (float)ci[a] * pg.cs32[a] - pg.slack32, 0.f);
					fp[a] = fmaxf((float)(ci[a] + 1) * pg.cs32[a] - c[a] - pg.slack32, 0.f);
				}
			}
			const float ext = DU ? en.extra32 : 0.f;
			const float amax = fminf(ai, pg.thr32) + ext;
#pragma unroll 1
			for (int s = 0; s < 27; s++) {
				const int r = s / 3, sub = s - 3 * r;
				const int oz = r / 3 - 1, oy = r - 3 * (r / 3) - 1;
				int nz = cz + oz, ny = cy + oy;
				float sx = 0.f, sy = 0.f, sz = 0.f;
				if (nz < 0) { nz += g.nc[2]; sz = -(float)g.box[2]; }
				if (nz >= g.nc[2]) { nz -= g.nc[2]; sz = (float)g.box[2]; }
				if (ny < 0) { ny += g.nc[1]; sy = -(float)g.box[1]; }
				if (ny >= g.nc[1]) { ny -= g.nc[1]; sy = (float)g.box[1]; }
				const int lz = nz - w2, ly = ny - w1;
				const float gy = oy == 0 ? 0.f : (oy > 0 ? fp[1] : fm[1]), gz = oz == 0 ? 0.f : (oz > 0 ? fp[2] : fm[2]);
				const float gyz = gy * gy + gz * gz;
				bool row_ok = gyz < amax && lz >= 0 && lz < d2 && ly >= 0 && ly < d1;
				const bool keep_lo = gyz + fm[0] * fm[0] < amax, keep_hi = gyz + fp[0] * fp[0] < amax;
				int xlo, xhi;
				if (sub == 0) {          // the unwrapped x range of a row shifted in y or z
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
				const int rowbase = fd0 * (ly + d1 * lz);
				const int jb = start[rowbase + xlo * xs], je = start[rowbase + (xhi + 1) * xs];
				const float qx = p32.x - sx, qy = p32.y - sy, qz = p32.z - sz;
				for (int j = jb; j < je; j++) {
					const float4 c = pos32[j];
					const float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
					if (__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx)) < fminf(ai, c.w) + ext) {
						const Particle pj = load_particle(pos + j);
						const D3 f = pair_force_term(i, pi, j, pj, g, nT, tab, 6 * nT * nT);
						ex += f.x; ey += f.y; ez += f.z;
						if (DU) du += pair_energy_term<2>(i, pi, j, pj, g, nT, en.uC, en.sx, en.sy, en.sz, false);
					}
				}
			}
		}
	}

	// =============================================================== hand the lists out again, longest first
	sm.part[0][tid] = ex; sm.part[1][tid] = ey; sm.part[2][tid] = ez;
	if (DU) s_dup[tid] = du;
	__syncthreads();   // (also: hist[] zeroed, tables staged, every warp's masks written)
	const int lbin = min(sm.cnt[tid], TILE_BINS - 1);
	atomicAdd(&sm.hist[lbin], 1);
	__syncthreads();
	if (tid < 32) {   // exclusive prefix over descending length
		constexpr int PER = (TILE_BINS + 31) / 32;
		int h[PER], sum = 0;
#pragma unroll
		for (int k = 0; k < PER; k++) { const int c = TILE_BINS - 1 - (tid * PER + k); h[k] = c >= 0 ? sm.hist[c] : 0; sum += h[k]; }
		int inc = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, inc, d); if (tid >= d) inc += v; }
		int run = inc - sum;
#pragma unroll
		for (int k = 0; k < PER; k++) { const int c = TILE_BINS - 1 - (tid * PER + k); if (c >= 0) sm.hist[c] = run; run += h[k]; }
	}
	__syncthreads();
	sm.order[atomicAdd(&sm.hist[lbin], 1)] = tid;
	__syncthreads();

	// =============================================================== phase 2: drain one list, FP64 (k_pair_force2's arithmetic)
	const int o = sm.order[tid];
	const int io = base + o;
	const bool act = io < N && !(g.slab && (gid[io] & GID_GHOST));
	if (!DU && !act && !pg.done) return;
	double ax = sm.part[0][o], ay = sm.part[1][o], az = sm.part[2][o];
	if (DU) du = s_dup[o];
	if (act) {
		const Particle po = load_particle(pos + io);
		const char *rowi = reinterpret_cast<const char *>(s_ptab + PTAB_STRIDE * po.type * nT);
		const char *rowu = reinterpret_cast<const char *>(s_utab + PTAB_STRIDE * po.type * nT);   // DU
		// Potential<T>, MD.h:895-930, on r^2 (x normal and positive)
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
		// Force<T>, MD.h:795-848: branch-free; sqrt and the division share one reciprocal square root, each finished with an
		// exact-residual correction.  true: the pair needs the general routine (asymmetric tables, r >= 2 rm)
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
			if (DU && !(in && !ok)) {
				const char *cu = rowu + (PTAB_STRIDE * 8) * pj.type;
				double uo = upot(in ? dr2 : 1.0, cu);
				uo = in ? uo : 0.0;
				double sxd = po.x * en.sx - pj.x * en.sx, syd = po.y * en.sy - pj.y * en.sy, szd = po.z * en.sz - pj.z * en.sz;
				double er2 = sxd * sxd + syd * syd + szd * szd;
				bool in2 = er2 < rc2 && j != io;
				double un = upot(in2 ? er2 : 1.0, cu);
				du += 0.5 * (uo - (in2 ? un : 0.0));   // the other half: the same pair in the neighbour's list
			}
			return in && !ok;
		};
		auto general = [&](int j, const Particle &pj) {
			D3 f = pair_force_term(io, po, j, pj, g, nT, tab, 6 * nT * nT);
			ax += f.x; ay += f.y; az += f.z;
			if (DU) du += pair_energy_term<2>(io, po, j, pj, g, nT, en.uC, en.sx, en.sy, en.sz, false);
		};
		const int hb = sm.hd[o];
		const int kwo = sm.kw[o];
		if (kwo == TILE_OVERFLOW) {
			int pre = 0;
			for (int r = 0; r < PAIR_NSEG; r++) {
				const int end = sm.g_end[r][hb], jb = sm.g_dlt[r][hb] + pre;
				for (int j = jb; j < jb + (end - pre); j++) {
					const Particle pj = load_particle(pos + j);
					if (fast(j, pj)) general(j, pj);
				}
				pre = end;
			}
		} else {
			// iterator over the set bits of the particle's mask words -> neighbour slot and record
			int k = -1, r = 0, lim = sm.g_end[0][hb], dl = sm.g_dlt[0][hb];
			unsigned m = 0u;
			auto fetch = [&](int &j) {
				while (m == 0u) { k++; m = sm.mask[k][o]; }
				const int c = (k << 5) + (__ffs(m) - 1);
				m &= m - 1u;
				while (c >= lim) { r++; lim = sm.g_end[r][hb]; dl = sm.g_dlt[r][hb]; }
				j = dl + c;
				return load_particle(pos + j);
			};
			int left = sm.cnt[o];
			if (left >= 2) {
				int j0, j1;
				Particle p0 = fetch(j0), p1 = fetch(j1);
				left -= 2;
				while (true) {
					int n0 = j0, n1 = j1;
					Particle q0 = p0, q1 = p1;
					const bool more = left >= 2;
					if (more) { q0 = fetch(n0); q1 = fetch(n1); left -= 2; }   // next two pairs: in flight during the math below
					const bool s0 = fast(j0, p0);
					const bool s1 = fast(j1, p1);
					if (s0 || s1) {
						if (s0) general(j0, p0);
						if (s1) general(j1, p1);
					}
					if (!more) break;
					j0 = n0; j1 = n1; p0 = q0; p1 = q1;
				}
			}
			if (left == 1) {
				int j0;
				const Particle p0 = fetch(j0);
				if (fast(j0, p0)) general(j0, p0);
			}
		}
	}
	if (DU) {   // the block's share of the dPotential, summed in particle order; then the force epilogue
		__syncthreads();
		sm.part[0][o] = act ? du : 0.0;
		__syncthreads();
		const double tot = block_sum(sm.part[0][tid]);
		if (tid == 0) en.partials[blockIdx.x] = tot;
		if (!act && !pg.done) return;
	}
	if (!act) {
		// (only reached with pg.done: every thread of the block takes part in the hand-over below)
	} else if (LANGEVIN) {
		const int id = lg.gid[io] & GID_MASK;
		double u[3];
		if (lg.ext_noise) {
			u[0] = lg.ext_noise[3 * id]; u[1] = lg.ext_noise[3 * id + 1]; u[2] = lg.ext_noise[3 * id + 2];
		} else {
			philox_uniform3(lg.seed, lg.step, (uint32_t)id, u);
		}
		const double lx = -lg.gamma * lg.vel[io] + lg.sigma * (2.0 * u[0] - 1.0);
		const double ly = -lg.gamma * lg.vel[cap + io] + lg.sigma * (2.0 * u[1] - 1.0);
		const double lz = -lg.gamma * lg.vel[2 * cap + io] + lg.sigma * (2.0 * u[2] - 1.0);
		acc[io] = lx + ax; acc[cap + io] = ly + ay; acc[2 * cap + io] = lz + az;
	} else {
		acc[io] += ax; acc[cap + io] += ay; acc[2 * cap + io] += az;
	}
	if (pg.done) {   // this block's accelerations are complete: release its seam block
		__threadfence();
		__syncthreads();
		if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(pg.done + bid), "r"(pg.epoch) : "memory");
	}
}

} // namespace smd
