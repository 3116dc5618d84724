/*
 * softmold_b200.h -- C ABI of the B200-native SoftMold MD timestep (libsoftmold_b200.so).
 *
 * The reference (LaradjiSoftMatter/SoftMold) has no plugin / FFI seam: its hot path is the set of C++ template
 * objects that `MD.cpp` instantiates and drives once per timestep.  This ABI replaces exactly those seams, one
 * entry point per reference interface (file:line relative to the reference root is cited on each), so a maintainer
 * can swap the objects for calls into this library (INTEGRATION.md shows the binding), and our own drop-in `MD`
 * host driver (softmold_b200/csrc/md_main.cpp) uses nothing else.
 *
 * Conventions
 *  - plain C: pointers + sizes, no C++ / torch types; every call returns int (SMD_OK == 0) and never throws;
 *    smd_last_error(ctx) returns the text of the last failure on that context (smd_last_error(NULL): create errors).
 *  - the caller owns every host buffer; the library owns all device memory; one context per GPU; calls on one
 *    context are not thread-safe (same as the reference objects, which borrow raw pointers into Blob).
 *  - host particle arrays are in the ORIGINAL particle order of the .mpd file: positions [n][3] doubles, types
 *    [n] int32, velocities / accelerations [n][3] doubles.  (Device storage is cell-sorted; the library permutes.)
 *  - all arithmetic is IEEE FP64 ("dtype f64"); cell and neighbour membership is bit-exact with the reference.
 *  - there is NO CPU fallback: every compute entry point fails with SMD_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SOFTMOLD_B200_H
#define SOFTMOLD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMD_ABI_VERSION 2

enum {
	SMD_OK = 0,
	SMD_ERR_ARG = 1,      /* bad argument / call order */
	SMD_ERR_CUDA = 2,     /* CUDA runtime failure or no usable device */
	SMD_ERR_CELL = 3,     /* particle outside the box / moved more than one cell per step / cell overflow
	                         (the reference: `throw 0` at cellOpt.h:541-552, :775-779, system.h:452-469) */
	SMD_ERR_IO = 4,       /* .mpd / output file problems */
	SMD_ERR_UNSUPPORTED = 5
};

/* molecule type ids, identical to the reference enum (include/algorithms/molecules.h:6-40) */
enum { SMD_MOL_BOND = 6, SMD_MOL_BEND = 7, SMD_MOL_CHAIN = 8, SMD_MOL_BEAD = 9, SMD_MOL_SOLID = 10, SMD_MOL_BOUNDARY = 11,
       SMD_MOL_RIGIDBEND = 12, SMD_MOL_PULLBEAD = 13, SMD_MOL_OFFSET_BOUNDARY = 14, SMD_MOL_FLOATING_BASE = 15,
       SMD_MOL_ZTORQUE = 16, SMD_MOL_ZPOWERPOTENTIAL = 17, SMD_MOL_NANOCORE = 18, SMD_MOL_BALL = 19 };

/* energy / force terms.  Bit masks select terms in smd_compute_forces; indices address out_terms[] arrays. */
enum { SMD_TERM_PAIR = 0, SMD_TERM_CHAIN = 1, SMD_TERM_BOND = 2, SMD_TERM_BEND = 3, SMD_TERM_BEAD = 4,
       SMD_TERM_BALL = 5,
       SMD_TERM_FIELD = 6,      /* BOUNDARY + FLOATING_BASE + ZTORQUE + ZPOWERPOTENTIAL: one-body / orientation fields */
       SMD_TERM_NANOCORE = 7, SMD_NTERMS = 8 };
#define SMD_MASK(term) (1u << (term))
#define SMD_MASK_LANGEVIN (1u << 16)
#define SMD_MASK_ALL_MOLECULES (SMD_MASK(SMD_TERM_CHAIN) | SMD_MASK(SMD_TERM_BOND) | SMD_MASK(SMD_TERM_BEND) | \
                                SMD_MASK(SMD_TERM_BEAD) | SMD_MASK(SMD_TERM_BALL) | SMD_MASK(SMD_TERM_FIELD) | \
                                SMD_MASK(SMD_TERM_NANOCORE))
#define SMD_MASK_ALL (SMD_MASK(SMD_TERM_PAIR) | SMD_MASK_ALL_MOLECULES | SMD_MASK_LANGEVIN)

/* Langevin noise source */
enum {
	SMD_NOISE_PHILOX = 0,   /* Philox4x32-10 keyed (seed, step, particle): the product default */
	SMD_NOISE_EXTERNAL = 1  /* uniforms supplied by the host with smd_set_noise (parity runs against the reference's
	                           MT19937 stream, algorithms/langevin.h:125-130) */
};

typedef struct smd_ctx smd_ctx;

/* Everything CellOpt / Verlet / Langevin take in their constructors:
 *   CellOpt::initialize  include/algorithms/cellOpt.h:153-227   (p, a, Fconst, Uconst, nParticles, nTypes, size, wrap, cutoff)
 *   Verlet ctor          include/algorithms/verlet.h:21-23       (p, a, v, nParticles, size, dt, wrap, aP)
 *   Langevin::initialize include/algorithms/langevin.h:110-135   (a, v, p, nParticles, gamma, dt, seed)            */
typedef struct smd_desc {
	int32_t abi_version;     /* SMD_ABI_VERSION */
	int32_t n_particles;
	int32_t n_types;
	int32_t device;          /* CUDA device ordinal */
	double box[3];           /* `size` */
	double cutoff;           /* `cutoff` (rc) */
	double dt;               /* `deltaT` */
	double gamma;            /* `gamma` */
	double temperature;      /* `initialTemp` */
	uint64_t seed;           /* `seed` */
	int32_t noise;           /* SMD_NOISE_* */
	int32_t track_unwrapped; /* keep the unwrapped copy aP used for diffusion (MD.cpp:96-105, verlet.h:318-330) */
	int32_t rank, nranks;    /* slab decomposition (0,1 for a single GPU) */
	int32_t reserved[8];
} smd_desc;

int smd_abi_version(void);
const char *smd_last_error(const smd_ctx *ctx);

/* number of CUDA devices this process can use; 0 (and SMD_ERR_CUDA from smd_create) when there is none */
int smd_device_count(void);

int smd_create(const smd_desc *desc, smd_ctx **out);
int smd_destroy(smd_ctx *ctx);

/* twoBodyFconst / twoBodyUconst, 6*nTypes^2 doubles each, row = type1*nTypes+type2
 * (Blob::getTwoBodyFconst/Uconst, include/system.h:256-257; layout include/potentials/laradjiRevalee.h:103-108,143-148) */
int smd_set_pair_tables(smd_ctx *ctx, const double *fC, const double *uC);

/* Blob::getPositions / getVelocities (include/system.h:247-249).  vel may be NULL (zeros). Resets the step state. */
int smd_set_particles(smd_ctx *ctx, const double *xyz, const int32_t *type, const double *vel);

/* molecule records, in the order of the .mpd file (the bead list of a BEAD molecule depends on that order,
 * include/system.h:2053-2070).  Constants as in MD.h:58-66:
 *   CHAIN  c[4] = {r0, kBond, cosTheta0, kBend}, blocks[n][3] = {start, nChains, length}   (system.h:1782-1866)
 *   BOND   c[2] = {r0, k},           ij[n][2]                                                 (system.h:1880-1934)
 *   BEND   c[2] = {cosTheta0, k},    ijk[n][3]                                                (system.h:1975-2040)
 *   BEAD   C[22*nTypes^2],           idx[n]                                                   (system.h:2043-2212)
 *   BALL   c[2] = {r0, k},           cj[n][2] = {centre, j}                                   (system.h:1936-1971)
 * and the remaining kinds of MD.cpp's switch (MD.cpp:414-478), none of which takes part in the box move (MD.cpp:642-669):
 *   BOUNDARY         c[4] = {dim, centre, unused, k},   idx[n]                                (system.h:2334-2348, MD.h:543-626)
 *   FLOATING_BASE    C[6*nTypes] (row = particle type), idx[n]                                (system.h:2402-2446, MD.h:457-494)
 *   ZTORQUE          c[4] = {a, b, c, d},               blocks[n][3] = {start, nChains, length}  (system.h:2582-2663, MD.h:1116-1140)
 *   ZPOWERPOTENTIAL  c[2] = {k, n},                     blocks[n][2] = {start, count}         (system.h:2665-2715, MD.h:1142-1153)
 *   NANOCORE         C[22*n] (one bead row per bead),   idx[n]                                (system.h:2215-2332, :3028-3075, :3708-3760)
 * SOLID, OFFSET_BOUNDARY, RIGIDBEND and PULLBEAD are parsed by the reference but do nothing in `MD` (default case of the
 * switch): smd_create_from_mpd accepts and skips them. */
int smd_add_chain(smd_ctx *ctx, int32_t n_blocks, const int32_t *blocks, const double c[4]);
int smd_add_bonds(smd_ctx *ctx, int32_t n, const int32_t *ij, const double c[2]);
int smd_add_bends(smd_ctx *ctx, int32_t n, const int32_t *ijk, const double c[2]);
int smd_add_beads(smd_ctx *ctx, int32_t n, const int32_t *idx, const double *C);
int smd_add_ball(smd_ctx *ctx, int32_t n, const int32_t *cj, const double c[2]);
int smd_add_boundary(smd_ctx *ctx, int32_t n, const int32_t *idx, const double c[4]);
int smd_add_floating_base(smd_ctx *ctx, int32_t n, const int32_t *idx, const double *C);
int smd_add_ztorque(smd_ctx *ctx, int32_t n_blocks, const int32_t *blocks, const double c[4]);
int smd_add_zpower(smd_ctx *ctx, int32_t n_blocks, const int32_t *blocks, const double c[2]);
int smd_add_nanocore(smd_ctx *ctx, int32_t n, const int32_t *idx, const double *C);
/* The three kinds only MDsubstrate.cpp's molecule switch evaluates (MDsubstrate.cpp:245-259 and :476-490; `MD` ignores them).
 * Forces only: neither dataExtraction::compute nor that driver's Metropolis sum has a term for them.
 *   OFFSET_BOUNDARY  doOffsetBoundaryForce system.h:2351-2363, offsetBoundaryF MD.h:502-527; idx[n]; c = {dim, centre, offset, k}
 *   RIGIDBEND        doRigidBendForce system.h:2366-2400, kmaxTorqueF MD.h:1073-1113; ij[n][2]; c = {zx, zy, zz, k, thetaD}
 *   PULLBEAD         doPullBeadForce system.h:2449-2486, harmonicFZ MD.h:434-447; idx[n]; c = {x0, y0, z0, k} */
int smd_add_offset_boundary(smd_ctx *ctx, int32_t n, const int32_t *idx, const double c[4]);
int smd_add_rigidbend(smd_ctx *ctx, int32_t n, const int32_t *ij, const double c[5]);
int smd_add_pullbead(smd_ctx *ctx, int32_t n, const int32_t *idx, const double c[4]);
/* a molecule the driver in charge parses and ignores (default case of MD.cpp:414-478: SOLID, OFFSET_BOUNDARY, RIGIDBEND,
 * PULLBEAD; of MDsubstrate.cpp:213-262: SOLID, BALL, FLOATING_BASE, ZTORQUE, ZPOWERPOTENTIAL, NANOCORE); registering it
 * keeps the molecule numbering of smd_observe equal to the file's */
int smd_add_inert(smd_ctx *ctx, int32_t kind);

/* Per-type friction, the `gammaType` command (MD.cpp:134-138, Langevin::compute algorithms/langevin.h:236-281).  What the
 * reference does with it, reproduced: the type lookup is commented out (`int type=0;//p[i].type;`), so EVERY particle
 * gets gamma_type[0]; and the noise amplitudes sT[] are computed once, at the temperature of the first evaluation, so a
 * later temperature ramp does not change them.  Replaces desc.gamma. */
int smd_set_gamma_type(smd_ctx *ctx, int32_t n_types, const double *gamma_type);

/* temperature may be ramped by the driver (MD.cpp:369-371) */
int smd_set_temperature(smd_ctx *ctx, double temperature);

/* uniforms in [0,1) for the NEXT Langevin evaluation, u[n][3] in original particle order, consumed once
 * (SMD_NOISE_EXTERNAL only).  Replaces MTRand::rand53 draws of Langevin::compute, algorithms/langevin.h:284-331. */
int smd_set_noise(smd_ctx *ctx, const double *u);

/* CellOpt::build (cellOpt.h:510-710): bin the current positions into the reference's cell grid and sort. */
int smd_build_cells(smd_ctx *ctx);

/* a = 0, then accumulate the selected terms on the current configuration:
 *   SMD_TERM_PAIR   CellOpt::build + computeForce            cellOpt.h:510-926, MD.h:795-848
 *   SMD_TERM_CHAIN..BALL   Blob::do*Force for every molecule of that kind   system.h:1782-2212
 *   SMD_MASK_LANGEVIN      Langevin::compute(temperature)     algorithms/langevin.h:232-331
 * `step` is the MD iteration index (Philox counter).  This is what MD.cpp:186-262 does before its loop. */
int smd_compute_forces(smd_ctx *ctx, uint32_t term_mask, int64_t step);

/* Finish an interrupted half kick after loading a checkpoint with initialTime != 0 (MD.cpp:274-308):
 * bead mass division + Verlet::second. */
int smd_resume(smd_ctx *ctx);

/* nsteps iterations i = first_step .. first_step+nsteps-1 of the loop MD.cpp:335-511 without store / measure /
 * box moves: bead mass division, Verlet::first, a=0, Langevin, build, pair + molecule forces, bead mass division,
 * Verlet::second.  Asynchronous: returns after enqueueing; any later call that reads results synchronises. */
int smd_step(smd_ctx *ctx, int64_t first_step, int32_t nsteps);

/* the two halves of one iteration, for drivers that store a checkpoint in the middle (MD.cpp:373-381) */
int smd_step_begin(smd_ctx *ctx, int64_t step);   /* bead mass division, Verlet::first, a = 0   (MD.cpp:340-366) */
int smd_step_end(smd_ctx *ctx, int64_t step);     /* Langevin ... Verlet::second                (MD.cpp:410-511) */

/* CellOpt::computePotential (cellOpt.h:928-1041) + Blob::do*Potential (system.h:2487-3167) as summed by
 * dataExtraction::compute (dataExtraction.h:839-990).  out_terms[SMD_NTERMS], indexed by SMD_TERM_*. */
int smd_potential(smd_ctx *ctx, double *out_terms);

/* Kinetic::compute, include/algorithms/dataCollection.h:666-674 */
int smd_kinetic(smd_ctx *ctx, double *out);

/* CellOpt::computeDPotential (cellOpt.h:1043-1180) + Blob::do*DPotential (system.h:3280-3757) for the proposed
 * component-wise scaling of the box (MD.cpp:608-675).  out_terms[SMD_NTERMS]. */
int smd_dpotential(smd_ctx *ctx, const double scale[3], double *out_terms);
/* the same sums left ON THE DEVICE: *d_terms = device pointer to SMD_NTERMS doubles, valid until the next energy call of this
 * context, written in stream order on smd_stream() -- nothing is copied and nothing waits.  For multi-GPU box moves: the
 * slab driver all-reduces this buffer in place (NCCL, ordered on the same stream) and reads the total back once, instead of
 * one device-to-host copy per rank followed by a host-staged all-reduce (MD.cpp:617-669 across ranks). */
int smd_dpotential_device(smd_ctx *ctx, const double scale[3], double **d_terms);

/* accepted box move: p *= scale, CellOpt::resize / Verlet::resize / setSize (MD.cpp:697-712, cellOpt.h:1525-1570) */
int smd_rescale(smd_ctx *ctx, const double scale[3], const double new_box[3]);

/* One Metropolis box-move trial exactly as MD.cpp:589-721 given the two rand53 draws the reference takes from
 * MTRand randNum(seed): u_fluct (:595) and u_accept (:688).  Returns accepted (0/1), the total dPotential incl.
 * tension*dA, and the box after the trial.  As in MD.cpp:657-666 the BALL and NANOCORE terms of smd_dpotential are
 * evaluated but NOT part of the Metropolis sum (the reference drops the return value), and the one-body fields have
 * no dPotential at all. */
int smd_mc_box_move(smd_ctx *ctx, double deltaLXY, double tension, double u_fluct, double u_accept,
                    int32_t *accepted, double *dU_total, double box_out[3]);

/* Arms the NEXT smd_step call: the pair kernel of its last step also sums the pair dPotential of the box scaling `scale`
 * in the same pass over the pairs as its forces (a Metropolis trial sees the positions of the last force evaluation:
 * Verlet::second only moves velocities, MD.cpp:511-615).  The first smd_dpotential call for that same scale afterwards,
 * with the particles untouched in between, takes its pair term from there.  A no-op where the fast path does not apply
 * (asymmetric tables, external noise): smd_dpotential then works as always.  Slab mode: every rank arms the same scale. */
int smd_arm_dpotential(smd_ctx *ctx, const double scale[3]);

/* smd_step(first_step, nsteps) followed by smd_mc_box_move(...) on the configuration it leaves -- the reference's cadence
 * (MD.cpp:335-729: resizeRate steps, then one trial), with the same results -- in one call, so that the pair kernel of the
 * LAST step can sum the pair dPotential of the proposed move in the same pass over the pairs as its forces (the trial sees
 * the positions of that force evaluation; Verlet::second only moves velocities).  Falls back to the two calls where that
 * does not apply (asymmetric tables, external noise, ...).  = smd_mc_propose + smd_arm_dpotential + smd_step + the trial. */
int smd_step_mc(smd_ctx *ctx, int64_t first_step, int32_t nsteps, double deltaLXY, double tension, double u_fluct,
                double u_accept, int32_t *accepted, double *dU_total, double box_out[3]);

/* ---- dataExtraction::compute's geometric observables, reduced on the device (dataExtraction.h:827-1693, default build: no
 * ANCHOR_DATA / FLAT_MEMBRANE / NANOPARTICLE blocks).  The reference walks the host arrays once per measureInterval; here the
 * particles never leave the GPU: one call = a few short kernels + one read-back of < 1 KB.
 *   SMD_OBS_BONDS    sums of the bond length over every BOND record (:861-893) and of cos(theta) and the two arm lengths
 *                    over every BEND record (:895-935), with the reference's inclusive image rule
 *   SMD_OBS_EXTENT   lo = min(size, min position), hi = max(0, max position) per axis: flicker = hi - lo (:1457-1487)
 *   SMD_OBS_KE_HIST  adds every particle to the kinetic-energy histogram, bin = int(0.5 v^2 / 0.0001) (:1511-1520); the
 *                    histogram lives on the device and accumulates over calls; smd_ke_histogram reads it
 *   SMD_OBS_MSD      msd_sum[k] = sum of |unwrapped - start|^2 over the particles of molecule k, msd_count[k] their number
 *                    (:1525-1663: BOND / BEND / BEAD records entry by entry, CHAIN blocks as index ranges, 0 for every
 *                    other kind), k in the order of the smd_add_* calls; needs track_unwrapped and smd_msd_start
 * Not in slab mode. */
#define SMD_OBS_BONDS 1u
#define SMD_OBS_EXTENT 2u
#define SMD_OBS_KE_HIST 4u
#define SMD_OBS_MSD 8u
typedef struct smd_observables {
	double lbond_sum;
	int64_t n_bond;
	double cos_bend_sum, lbend_sum[2];
	int64_t n_bend;
	double lo[3], hi[3];
	int32_t n_molecules;     /* entries written to msd_sum / msd_count */
} smd_observables;
int smd_observe(smd_ctx *ctx, uint32_t what, smd_observables *out, double *msd_sum, int64_t *msd_count, int32_t msd_cap);
/* aPStart of the reference (MD.cpp:96-105, dataExtraction.h:790-803): remember the unwrapped positions of this moment */
int smd_msd_start(smd_ctx *ctx);
/* kEnergyDensity (dataExtraction.h:1680-1693): *n_bins = highest populated bin + 1; counts[0 .. min(cap, *n_bins)) if not NULL */
int smd_ke_histogram(smd_ctx *ctx, int64_t *counts, int64_t cap, int64_t *n_bins);

/* read back (original particle order).  Any pointer may be NULL. */
int smd_get_particles(smd_ctx *ctx, double *xyz, int32_t *type, double *vel);
int smd_get_forces(smd_ctx *ctx, double *acc);
int smd_get_unwrapped(smd_ctx *ctx, double *xyz);
int smd_get_box(smd_ctx *ctx, double box[3]);

/* Asynchronous read-back for the checkpoint / trajectory / observable writers (MD.cpp:373-381 Script::write +
 * xyzFormat::store, dataExtraction::compute :525-543 -- all synchronous in the reference, where one 240 000-particle
 * store is cheap next to 1 000 CPU steps; next to 1 000 GPU steps it is not, so the driver moves the text formatting
 * to a worker thread and the device keeps stepping).
 *   smd_host_alloc / smd_host_free   page-locked host memory (so that the copy really is asynchronous)
 *   smd_snapshot        gathers the state into original particle order on the context's stream and starts the
 *                       device-to-host copies on a separate copy stream; returns a ticket at once.  Any of xyz [n][3],
 *                       vel [n][3], unwrapped [n][3] may be NULL; the buffers must stay valid until the ticket was
 *                       waited for.  At most two tickets may be outstanding.
 *   smd_snapshot_wait   blocks the calling thread -- ANY thread -- until the copies of that ticket have landed */
int smd_host_alloc(void **ptr, size_t bytes);
int smd_host_free(void *ptr);
int smd_snapshot(smd_ctx *ctx, double *xyz, double *vel, double *unwrapped, int64_t *ticket);
int smd_snapshot_wait(smd_ctx *ctx, int64_t ticket);

/* cell membership as the reference computes it (cellOpt.h:530-556): key of every particle, and the particles of
 * every occupied cell in the reference's list order (descending particle index, cellOpt.h:572-585).
 * n_cells_xyz[3] = nCells.  cell_key / cell_rank are [n]: rank = position of the particle in its cell's list. */
int smd_get_cell_ids(smd_ctx *ctx, int32_t n_cells_xyz[3], int32_t *cell_key, int32_t *cell_rank);

/* number of in-range pairs (r^2 < rc^2), and per-particle neighbour counts [n] (may be NULL) */
int smd_count_pairs(smd_ctx *ctx, int64_t *total, int32_t *per_particle);

/* wait for all enqueued work; reports deferred device-side errors (SMD_ERR_CELL) */
int smd_synchronize(smd_ctx *ctx);

/* device pointers of the resident state for zero-copy interop (torch / NCCL plumbing): sorted order!
 *   which: 0 positions {x,y,z,type-bits}[n] (32 B records), 1 velocities SoA [3][cap], 2 accelerations SoA [3][cap],
 *          3 original index of each slot int32[n] */
int smd_device_ptr(smd_ctx *ctx, int32_t which, void **ptr, size_t *bytes);

/* CUDA stream (cudaStream_t) all work of this context is enqueued on, for event timing by the caller */
int smd_stream(smd_ctx *ctx, void **stream);

/* counters: kernels launched since creation, cell rebuilds */
int smd_stats(smd_ctx *ctx, int64_t *kernel_launches, int64_t *rebuilds);

/* Per-phase device timing with CUDA events recorded on the context's stream around the kernels of each phase of
 * the timestep (the reference only prints wall seconds per measure interval, MD.cpp:527-532).
 * smd_profile(ctx, mask): bit p enables phase p; 0 disables; accumulators are reset.
 * smd_profile_read synchronises and returns accumulated milliseconds and bracket counts per phase. */
enum { SMD_PHASE_INTEGRATE1 = 0,  /* bead mass, Verlet::first (+ cell tagging), a = 0   MD.cpp:340-366 */
       SMD_PHASE_BUILD = 1,       /* CellOpt::build                                     MD.cpp:412     */
       SMD_PHASE_PAIR = 2,        /* CellOpt::computeForce                              MD.cpp:413     */
       SMD_PHASE_MOLECULES = 3,   /* Blob::do*Force                                     MD.cpp:414-509 */
       SMD_PHASE_LANGEVIN = 4,    /* Langevin::compute                                  MD.cpp:410     */
       SMD_PHASE_INTEGRATE2 = 5,  /* Verlet::second                                     MD.cpp:511     */
       SMD_PHASE_STEP = 6,        /* one whole smd_step_begin + smd_step_end                           */
       SMD_PHASE_EXCHANGE = 7,    /* slab mode: migration + halo pack / unpack                         */
       SMD_PHASE_FUSED = 8,       /* CHAIN-only systems: chain forces + Verlet::second + next Verlet::first in one
                                     kernel (then phases 0, 3, 5 only count the first / last step of a batch) */
       SMD_PHASE_PAIR_DU = 9,     /* the pair kernel of a step that also sums the dPotential of the box move (smd_step_mc) */
       SMD_PHASE_BUILD_HIST = 10, /* the kernels of the cell build one by one (inside SMD_PHASE_BUILD): histogram,        */
       SMD_PHASE_BUILD_SCAN = 11, /*   exclusive scan of the offset table,                                                */
       SMD_PHASE_BUILD_PLACE = 12,/*   claim of a position inside the cell's range,                                       */
       SMD_PHASE_BUILD_REORDER = 13, /* rank inside the cell (descending index) + move of the records                     */
       SMD_NPHASES = 16 };
int smd_profile(smd_ctx *ctx, uint32_t phase_mask);
int smd_profile_read(smd_ctx *ctx, double ms[SMD_NPHASES], int64_t count[SMD_NPHASES]);

/* The step on the device's own clock.  With programmatic dependent launches the kernels of a step overlap -- a block of the
 * next kernel becomes resident as soon as one of this kernel's leaves --, which event pairs cannot show (and switch off).
 * smd_timeline(ctx, 1): every smd_step call of >= 3 steps stamps %globaltimer in the kernels of its second-to-last step;
 * smd_timeline_read: us15[3 k + {0, 1, 2}] = microseconds after the first block of the scan at which kernel k (0 k_scan,
 * 1 k_place, 2 k_reorder, 3 k_pair_force2, 4 k_chain_kick) saw its first block start working, its last block start, and its
 * last block end; -1: that kernel did not run.  Off by default (costs one cached load per block when off). */
int smd_timeline(smd_ctx *ctx, int32_t enable);
int smd_timeline_read(smd_ctx *ctx, double *us15);

/* FP64 pipe peak of the context's device in TFLOP/s, measured with a dependency-chain micro-kernel: with fused
 * multiply-add, and with separate multiply + add (what this library issues: it is compiled without contraction to
 * stay bit-exact with the reference's x86-64 build).  Roofline denominator for the pair kernel. */
int smd_fp64_peak(smd_ctx *ctx, double *fma_tflops, double *muladd_tflops);

/* ------------------------------------------------------------------ slab decomposition over several GPUs
 * The reference is one shared-memory process (its MPI attempt, MPImem.h / runMPI.cpp, does not compile any more);
 * large flat-bilayer systems are decomposed here into slabs of cell columns along x, one context (= one GPU, one
 * process) per slab, SURVEY.md 8(e).  A context created with desc.nranks > 1 is a slab rank:
 *   - desc.n_particles is the GLOBAL particle count, desc.box the global box; desc.reserved[0] = local capacity in
 *     particles (0: 1.25 * n/nranks + halo allowance), desc.reserved[1] = entries per halo message (0: default);
 *   - smd_set_particles takes the GLOBAL arrays (every rank passes the same data) and keeps what this rank owns plus
 *     its ghost columns; molecule records use global particle indices; CHAIN, BOND and BEND molecules are supported (the
 *     members of a record must lie within SMD_SLAB_HALO cell columns of each other, else the run stops with SMD_ERR_CELL);
 *   - every step the ranks exchange migrating particles and the ghost-column halo (SMD_SLAB_HALO cell columns) by
 *     writing straight into the neighbour's receive buffer through peer memory (NVLink), inside smd_step_begin
 *     (send) and smd_step_end (receive): wire the buffers once with smd_slab_ipc_handle / smd_slab_connect_ipc
 *     (neighbour in another process) or smd_slab_recv_buffer / smd_slab_connect_ptr (same process);
 *   - smd_potential / smd_dpotential / smd_kinetic / smd_count_pairs return this rank's PARTIAL sums: every pair and
 *     bonded term is counted by exactly one rank, the caller all-reduces (smd_count_pairs: sum of the owned
 *     particles' neighbour counts, NOT halved);  smd_mc_box_move is replaced by smd_mc_propose -> smd_dpotential ->
 *     all-reduce -> smd_mc_accept -> smd_rescale, evaluated identically on every rank;
 *   - read-back is per rank: smd_slab_get_local. */
#define SMD_SLAB_HALO 2

/* owned cell columns [col_lo, col_hi) of `rank` when n_cols columns are dealt to nranks slabs (pure host function) */
int smd_slab_columns(int32_t n_cols, int32_t nranks, int32_t rank, int32_t *col_lo, int32_t *col_hi);
/* which of the n particles a slab rank keeps: flags[i] = 0 none, 1 owned, 2 ghost (pure host function; the cell
 * column is int(x / (Lx / int(Lx / cutoff))) with the upper-edge clamp, exactly CellOpt::build's, cellOpt.h:530-539) */
int smd_slab_select(const double box[3], double cutoff, int32_t nranks, int32_t rank, int32_t n, const double *xyz,
                    int32_t *flags);

/* own receive buffer of `side` (0: filled by the left neighbour, 1: by the right one) */
int smd_slab_recv_buffer(smd_ctx *ctx, int32_t side, void **ptr, size_t *bytes);
/* 64-byte CUDA IPC handle of that buffer, to be shipped to the neighbour process */
int smd_slab_ipc_handle(smd_ctx *ctx, int32_t side, void *handle64);
/* dir 0: the buffer is the LEFT neighbour's side-1 receive buffer; dir 1: the RIGHT neighbour's side-0 one */
int smd_slab_connect_ipc(smd_ctx *ctx, int32_t dir, const void *handle64);
int smd_slab_connect_ptr(smd_ctx *ctx, int32_t dir, void *peer_buffer);
/* the two halves of the exchange, for callers that move particles themselves (smd_step_begin / _end call them) */
int smd_slab_exchange_send(smd_ctx *ctx);
int smd_slab_exchange_recv(smd_ctx *ctx);
/* Load THIS rank's own particles only (restart from per-rank data): n particles with their global indices; nothing
 * else is kept.  Collective in effect: every rank calls it, then smd_compute_forces -- the ghost columns arrive from
 * the neighbours through the ordinary exchange (sent here, received by the force evaluation), and a particle that
 * sits up to SMD_SLAB_HALO columns outside this rank's range migrates to its owner on the way. */
int smd_slab_set_local(smd_ctx *ctx, int32_t n, const int32_t *gid, const double *xyz, const int32_t *type, const double *vel);
/* live counts of this rank (synchronises) */
int smd_slab_counts(smd_ctx *ctx, int32_t *n_local, int32_t *n_owned);
/* owned particles of this rank, arbitrary order: global index, position, type, velocity, acceleration.  The arrays
 * must hold the local capacity (smd_slab_capacity); any pointer may be NULL.  *n = number returned. */
int smd_slab_capacity(smd_ctx *ctx, int32_t *capacity);
int smd_slab_get_local(smd_ctx *ctx, int32_t *n, int32_t *gid, double *xyz, int32_t *type, double *vel, double *acc);

/* The host arithmetic of one Metropolis box-move trial, MD.cpp:591-613 and :677-695, shared by smd_mc_box_move and by
 * multi-rank drivers: propose new box + component-wise scale from u_fluct; decide from the all-reduced sum of the
 * dPotential terms.  Returns 1 = accepted, 0 = rejected in *accepted. */
int smd_mc_propose(const double box[3], double deltaLXY, double u_fluct, double new_box[3], double scale[3]);
int smd_mc_accept(double dU_terms_sum, double tension, const double box[3], const double new_box[3], double temperature,
                  double u_accept, int32_t *accepted, double *dU_total);

/* ------------------------------------------------------------------ host-side file boundary (no GPU needed)
 * `.mpd` reader / writer with the semantics of Script<T,Blob>::read/write + Blob::input/output
 * (include/fileFormats/scriptFormat.h:64-95, include/system.h:589-1749). */
typedef struct smd_mpd smd_mpd;

int smd_mpd_read(const char *name_without_ext, smd_mpd **out, char *err, size_t errlen);
int smd_mpd_write(const smd_mpd *m, const char *name_without_ext, char *err, size_t errlen);
void smd_mpd_free(smd_mpd *m);

/* scalar commands by their .mpd name ("gamma", "initialTemp", "seed", "nTypes", "deltaLXY", "tension", ...);
 * present = 0 when the command was not in the file (then *value is the reference default) */
int smd_mpd_get_scalar(const smd_mpd *m, const char *command, double *value, int32_t *present);
int smd_mpd_set_scalar(smd_mpd *m, const char *command, double value);
int smd_mpd_get_size(const smd_mpd *m, double size[3]);
int smd_mpd_set_size(smd_mpd *m, const double size[3]);
/* borrowed pointers into the object, valid until smd_mpd_free: xyz [n][3], type [n], vel [n][3] */
int smd_mpd_particles(smd_mpd *m, int32_t *n, double **xyz, int32_t **type, double **vel);
int smd_mpd_pair_tables(smd_mpd *m, int32_t *n_types, double **fC, double **uC);
int smd_mpd_n_molecules(const smd_mpd *m);
/* molecule k: type, number of bond records, ints per record, borrowed pointers to records and constants */
int smd_mpd_molecule(smd_mpd *m, int32_t k, int32_t *type, int32_t *n_records, int32_t *record_width,
                     int32_t **records, int32_t *n_constants, double **constants);

/* convenience: create a context from a parsed file (tables, particles, molecules all set) */
int smd_create_from_mpd(smd_mpd *m, int32_t device, int32_t noise, int32_t track_unwrapped, smd_ctx **out);
/* the same with the molecule switch of a named driver of the reference:
 *   SMD_DRIVER_MD         MD.cpp:414-478 (and MDanneal.cpp, same switch): BALL, FLOATING_BASE, ZTORQUE, ZPOWERPOTENTIAL, NANOCORE
 *                         act; OFFSET_BOUNDARY, RIGIDBEND, PULLBEAD are parsed and ignored
 *   SMD_DRIVER_SUBSTRATE  MDsubstrate.cpp:213-262: the other way round (and no NANOCORE mass division)
 * CHAIN, BOND, BEND, BEAD, BOUNDARY act under both; SOLID under neither. */
#define SMD_DRIVER_MD 0
#define SMD_DRIVER_SUBSTRATE 1
int smd_create_from_mpd_driver(smd_mpd *m, int32_t device, int32_t noise, int32_t track_unwrapped, int32_t driver, smd_ctx **out);

#ifdef __cplusplus
}
#endif
#endif
