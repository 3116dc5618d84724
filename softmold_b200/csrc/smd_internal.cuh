// Internal definitions shared by the CUDA translation units of libsoftmold_b200.so.
// Compiled for sm_100a only, with --fmad=false: the reference is g++ -O3 on baseline x86-64 (no FMA), and cell /
// neighbour membership has to be bit-exact with it, so no multiply-add may be contracted anywhere in this library.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/softmold_b200.h"

namespace smd {

// One particle record, 32 B, 32 B aligned: the device twin of the reference's position<T> {x,y,z,int type}
// (include/algorithms/dataTypes.h:183-351).  The 4 spare bytes carry the particle's cell coordinates in the
// reference grid, packed 11/11/10 bits (x,y,z), filled by the binning kernel.
struct __align__(32) Particle {
	double x, y, z;
	int type;
	unsigned cell;
};

__host__ __device__ inline unsigned pack_cell(int cx, int cy, int cz) { return (unsigned)cx | ((unsigned)cy << 11) | ((unsigned)cz << 22); }
__host__ __device__ inline void unpack_cell(unsigned c, int &cx, int &cy, int &cz) { cx = c & 2047u; cy = (c >> 11) & 2047u; cz = c >> 22; }

// geometry passed by value to kernels (changes only through host-driven box moves)
struct Geom {
	double box[3];
	double cs[3];   // cell size  = box / nc          (cellOpt.h:207-209)
	int nc[3];      // cells per axis = int(box / rc) (cellOpt.h:194-196)
	double rc2;     // cutoff^2
	// slab decomposition along x (multi-GPU): this context owns the cell columns [col_lo, col_hi) of the global grid
	// and keeps `halo` ghost columns on each side (periodic: column indices wrap modulo nc[0]).  slab == 0: whole box.
	int slab, col_lo, col_hi, halo;
	// x sub-cells: the SORT KEY (and the offset table start[]) splits every reference cell into xs slices along x, so a
	// stencil row -- the three cells of a (y,z) row are contiguous in the sorted order -- is sorted by x to a resolution
	// of cs[0] / xs and k_pair_force2 reads only the slices within the particle's reach (4 sigma of the row's 6).
	// Particle::cell still holds the reference's own cell coordinates.  xs = 1: asymmetric tables.
	int xs;
	double finv;    // xs / cs[0]
};

// particle count of a launch: a host value (single GPU: N never changes) or a device word (slab mode: the number of
// local particles changes every step through migration and the halo, and the host never waits for it)
struct Cnt {
	int n;
	const int *dn;
	__device__ __forceinline__ int get() const { return dn ? *dn : n; }
};

// slab mode: gid[] carries the ghost flag of a slot in bit 30 (ghost = copy of a particle another rank owns);
// a record whose cell word is CELL_DEAD is dropped by the next build
constexpr int GID_GHOST = 1 << 30;
constexpr int GID_MASK = GID_GHOST - 1;
constexpr unsigned CELL_DEAD = 0xffffffffu;

// device-resident active window of the cell grid: the bounding box of occupied cells plus one cell each side.
// win[0..2] origin, win[3..5] extent, win[6] number of cells in the window, win[8..9] the resolution of the 16-bit
// window-relative coordinates of pos16[] and its inverse (float bit patterns; see quant_res).
enum { WIN_ORG = 0, WIN_DIM = 3, WIN_NCELLS = 6, WIN_FD0 = 7, WIN_RES = 8, WIN_INVRES = 9, WIN_DIRTY = 10, WIN_WORDS = 12 };

enum { ERR_OUT_OF_BOX = 1, ERR_WINDOW_CAP = 4, ERR_SLAB_MIGRATION = 8, ERR_SLAB_MSG_CAP = 16, ERR_SLAB_CAPACITY = 32,
       ERR_SLAB_TIMEOUT = 64, ERR_SLAB_MISSING = 128 };

// The tagging pass (the kernel that moves the particles) may fill the histogram of the NEXT build's counting sort itself
// (bin_particle): the window it bins into, the histogram, the key of every slot.  count == nullptr: k_bin does it later.
struct BinArgs { int *win; int *count; int *cellOfSlot; };

// slab halo / migration messages.  One receive buffer per side (0: from the left neighbour, 1: from the right one)
// and parity of the exchange sequence number; the SENDER writes it directly through peer memory (NVLink P2P, or the
// same device in the single-process test harness) and publishes {count, seq} in the header last.
struct __align__(32) SlabMsgHeader { int n, seq; int pad[6]; };
struct __align__(32) SlabMsgEntry {    // 96 B
	double x, y, z; int type; unsigned cell;     // the Particle record
	double vx, vy, vz; int gid; int pad0;        // gid carries GID_GHOST for halo copies; velocities only for migrants
	double ux, uy, uz; double pad1;              // unwrapped position (migrants, when tracked)
};
struct SlabComm {
	char *send[2];          // peer pointers: the left / right neighbour's message buffer that faces this rank (parity 0; parity 1 follows)
	char *recv[2];          // own message buffers: [0] shared with the left neighbour, [1] with the right one
	int *counters;          // [0..1] entries claimed per direction, [2] blocks done
	size_t parity_stride;   // bytes between the two parity copies of a buffer
	int capmsg;             // entries per message
	// Who crosses the link.  pull = 0 (default): the sender writes through peer memory into the neighbour's buffer send[d];
	// every writing thread needs a system-scope fence before its block may be counted, and since the boundary columns are
	// spread over every block of a cell-sorted array, every block of the step seam pays that NVLink round trip.
	// pull = 1 (SMD_SLAB_PULL=1): the sender packs into ITS OWN buffer recv[d] -- local stores, a device-scope fence -- and the
	// receiver's unpack kernel reads the neighbour's buffer send[side] through the link after acquiring its header at system
	// scope.  Measured on 2 x B200 (C5 tile, profiles/r02h_slab_transport.md): the seam gets 11.6 us cheaper (96.1 -> 84.5),
	// the unpack 14.1 us dearer (16.0 -> 30.1: remote reads on the critical path of the step): 860.8 -> 870.0 us per step.
	// Same messages, bit-identical results; push stays.
	int pull;
	int gpu_fence;          // EXPERIMENT (SMD_SLAB_FENCE_GPU=1): writers of the push transport fence at device scope only, see slab_publish
};

struct PairGeo {            // phase-1 constants of k_pair_force2
	float thr32;            // rc^2 + FP32 rounding margin
	float margin32;         // that margin
	float slack32;          // FP32 rounding of a coordinate difference against a cell face
	float cs32[3];          // cell size
	float rmin32;           // smallest positive per-type phase-1 radius (pos16 path: widening of the energy modes' cutoffs)
	float finv32;           // x slices per unit length (Geom::finv)
	int *done;              // completion words, one per 32 slots, for a programmatic dependent launch of the step seam (else null)
	int epoch;              // value a block stores there
	int first;              // first slot of this launch (the tail launch of a hybrid pair of launches starts past the main one's)
	int part0;              // index of this launch's first block sum in EnergyArgs::partials
	int nowait;             // tail launch of a hybrid pair: does not wait for the main launch, which it does not depend on (see k_pair_force2)
};

// energy modes of k_pair_force2: the proposed scaling, the widening of the phase-1 cutoffs, where the block sums go;
// uC / utab (the potential's tables) only for EMODE 3, which runs on the force tables
struct EnergyArgs { double sx, sy, sz; float extra32; double *partials; const double *uC, *utab; };

struct ChainBlock { int start, nChains, len; double c[4]; };
struct BondList { int n; int *d_ij; double c[2]; };
struct BendList { int n; int *d_ijk; double c[2]; };
struct BallList { int n; int *d_cj; double c[2]; };
// BOUNDARY / FLOATING_BASE / ZTORQUE / ZPOWERPOTENTIAL / NANOCORE (kind = the reference's molecule type id)
struct FieldMol {
	int kind, n;               // n records (particles or blocks)
	int *d_idx;                // [n] original particle indices (BOUNDARY, FLOATING_BASE, NANOCORE)
	double *d_C;               // device constants (FLOATING_BASE 6*nT, NANOCORE 22*n)
	double c[4];               // host constants (BOUNDARY, ZTORQUE, ZPOWERPOTENTIAL)
	std::vector<int> blocks;   // host block records (ZTORQUE [n][3], ZPOWERPOTENTIAL [n][2])
	std::vector<int> host_idx;          // NANOCORE: host copy of the bead indices and radii (mass division in the fused seam)
	std::vector<double> host_radius;
};
struct BeadMol {
	int nOwn, nAll;        // beads of this molecule / of the list assembled from molecules j >= i (system.h:2053-2070)
	int *d_beads;          // [nAll] original particle indices, own beads first
	double *d_C;           // [22*nT*nT]
	double radius;         // C[BEADRADIUS=4]
	int mol_index;         // position in the molecule list (file order)
	std::vector<int> own;  // host copy of own bead indices
};

} // namespace smd

struct smd_ctx {
	smd_desc desc;
	int N, nT, cap;
	smd::Geom geom;
	double temperature;
	std::string err;
	cudaStream_t stream;
	int device;

	// resident particle state, cell-sorted ("slot" order); double-buffered for the per-step reorder
	smd::Particle *pos[2];
	float4 *pos32;    // FP32 mirror {x,y,z,type} of pos[cur], written by the reorder (phase 1 of the pair force kernel)
	float *acut;      // [nT] FP32 phase-1 class cutoff per type, margin included (see k_pair_force2)
	uint2 *pos16;     // 8-byte phase-1 candidates {x,y,z: 16-bit window-relative fixed point; cutoff^2 as a bf16} of pos[cur]
	float *arad;      // [nT] phase-1 class radius per type (rc, rm, or -1: interacts with nothing), no margin
	double *ptab;     // [nT*nT][PTAB_STRIDE] padded force table + exact branch thresholds
	double *utab;     // same layout, potential constants (energy modes of the two-phase kernel)
	bool force_onephase_energy = false;   // SMD_ENERGY_ONEPHASE=1: use the one-phase half-stencil energy kernels (A/B checks)
	std::vector<float> acut_raw;   // host: rm^2 / +inf / -1 per type, before the margin
	std::vector<float> acut_host;  // host copy of acut[] (source of an un-waited upload)
	smd::PairGeo pgeo; // rc^2 + FP32 rounding margin etc. for that phase
	double *vel[2];   // SoA [3][cap]
	double *unw[2];   // SoA [3][cap] unwrapped positions (optional)
	int *gid[2];      // original index of slot
	int cur;          // current buffer of vel / unw / gid
	int pcur;         // current buffer of pos (flips on its own when the fused step kernel writes the drifted positions)
	// smd_step_mc: the last step's pair kernel also sums the dPotential of the box move that follows (k_pair_force2 EMODE 3)
	bool du_for_last = false, du_armed = false, du_ready = false, no_du_fuse = false;
	bool no_seam_pack = false;
	bool pdl = true;            // step seam as a programmatic dependent of the pair kernel (SMD_PDL=0: plain stream order)
	bool pdl_chain = false;     // SMD_PDL=2: the build kernels and the pair kernel are programmatic dependents too
	int *pair_done = nullptr;   // [blocks] completion words, see PairGeo::done
	int pair_epoch = 0;
	smd::EnergyArgs du_en;
	double *du_partials = nullptr;   // block sums of the armed dPotential (their own buffer: nothing else writes it)
	size_t du_partials_n = 0;
	int sm_count = 148;      // SMs of the device
	int pair_tail = 0;       // SMD_PAIR_TAIL=1 (experiment): hand the last, partial round of pair blocks to the three-thread engine
	int du_nparts = 0;       // block sums the armed force + dPotential launch(es) leave in du_partials
	bool timeline = false;   // smd_timeline: stamp the kernels of the second-to-last step of every smd_step call
	int pair3 = -1;         // SMD_PAIR3: the three-threads-per-particle pair engine: 0 never, 1 always, default: systems of at most pair3_max particles
	int pair3_max = 0;
	bool no_fuse = false;   // SMD_NO_FUSE=1: always run the separate chain / Verlet kernels (A/B checks)
	double *acc;      // SoA [3][cap]
	double *acc2;     // alternate buffer for builds that must carry live accelerations along
	bool acc_live;    // acc holds forces a later kick still needs
	int *slot_of;     // [N] slot of original index

	// cell grid
	long long cellcap = 0;
	long long cellcap_limit = 64ll << 20;   // the offset tables never grow beyond this many entries
	int xs_wanted;     // x slices per cell when the fast path applies (SMD_XSUB, default 4)
	int *count, *start, *cursor;
	unsigned long long *scan_state;   // k_scan: [SCAN_BLOCKS] prefix words, then [SCAN_BLOCKS] chunk totals
	unsigned *scan_barrier = nullptr; // grid-barrier counter of k_scan's re-binning fallback
	int *cellOfSlot;
	int2 *order;      // {previous slot, original index} of every position claimed by k_place
	int *win;         // [2][WIN_WORDS] window descriptors (device): [wcur] of the current sorted order, [wcur ^ 1] the one the next
	                  // tagging pass bins into (published by every build: occupied box + 1 cell)
	int wcur = 0;
	bool no_slab_prebin = false;   // SMD_NO_SLAB_PREBIN=1: slab mode keeps the histogram pass of its own (k_bin) in every build (A/B)
	bool prebin = true;         // SMD_NO_PREBIN=1: the histogram of the build is always a pass of its own (k_bin)
	bool hist_pending = false;  // count[] holds the histogram of the current positions under win[wcur ^ 1] (a tagging pass filled it)
	bool next_win_valid = false; // win[wcur ^ 1] was published by the last build and nothing has changed the geometry since
	int *bbox;        // [6] min xyz, max xyz accumulators
	int *errflag;
	bool cells_valid;  // sorted order + start[] describe the current positions

	// pair tables
	double *fC, *uC;
	bool tables_set, particles_set;
	bool tables_symmetric;   // fC and uC rows (t1,t2) == (t2,t1): enables the orientation-free fast path of the pair kernel

	// molecules
	std::vector<smd::ChainBlock> chains;
	std::vector<smd::BondList> bonds;
	std::vector<smd::BendList> bends;
	std::vector<smd::BallList> balls;
	std::vector<smd::BeadMol> beads;
	std::vector<smd::FieldMol> fields;
	int n_molecules;
	struct MolRef { int kind, first, count; };   // molecule k of the smd_add_* order: kind + its entries in the vector of that kind
	std::vector<MolRef> mol_order;

	// device-side observables (smd_observe)
	unsigned long long *obs_buf = nullptr;   // [0..5] extent keys, [6] overflow count, [8 ..] sums
	unsigned long long *obs_host = nullptr;  // pinned mirror
	int obs_words = 0;
	unsigned long long *ke_bins = nullptr, *ke_overflow = nullptr;
	std::vector<std::pair<long long, long long>> ke_spill;   // (bin, count) beyond the device table
	double *msd_start = nullptr;             // [3][N] by original index

	// noise
	bool sigma_frozen = false;          // per-type friction path: noise amplitude fixed at its first temperature
	double sigma_temperature = -1.0;
	double *noise;     // [N][3] original order, external uniforms
	bool noise_ready;

	// scratch
	double *partials;  // block partial sums
	double *scalars;   // small device result buffer
	int *icount;       // per-particle int scratch
	double *stage;     // [N][3] staging in original order
	int *istage;       // [N]
	double *h_pinned;  // pinned host scratch (scalars)
	double *terms_dev = nullptr;   // [SMD_NTERMS] result of smd_dpotential_device
	int *export_i = nullptr;       // smd_slab_get_local: compacted gid / type [2][cap]
	double *export_d = nullptr;    //                     compacted xyz / vel / acc [9][cap]
	int *import_bad = nullptr;     // first out-of-box / out-of-range particle found by the import kernel (smd_set_particles)

	long long launches, rebuilds;

	// slab decomposition (desc.nranks > 1): N is then only the launch bound (= cap); the live counts sit on the device
	bool slab = false;
	int n_global = 0;      // particles of the whole system (size of slot_of[])
	int *dN = nullptr;     // [0] local particles (owned + ghost) of the current sorted order, [1] count after an unpack
	smd::SlabComm comm;    // message buffers (own receive side + the neighbours' as peer pointers)
	char *recv_base[2] = {nullptr, nullptr};
	size_t recv_bytes = 0;
	void *ipc_opened[2] = {nullptr, nullptr};
	bool peer_set[2] = {false, false};
	int xseq = 0;          // exchange sequence number (identical on every rank)
	bool exch_pending = false;   // a pack was issued, the matching unpack not yet
	bool ext_valid = false;      // dN[1] holds the count after an unpack (consumed by the next build)
	int *d_export_counter = nullptr;

	// asynchronous snapshots (smd_snapshot): gather buffer, copy stream, one event per ticket parity
	double *snap_stage = nullptr;          // [9][cap]: xyz, vel, unwrapped in original order
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t snap_gathered = nullptr, snap_done[2] = {nullptr, nullptr};
	std::atomic<long long> snap_seq{0};   // (smd_snapshot_wait reads it from the writer thread)

	// per-phase event timing (smd_profile)
	struct ProfSpan { int phase; cudaEvent_t e0, e1; };
	uint32_t prof_mask = 0;
	std::vector<ProfSpan> prof_pending;
	std::vector<cudaEvent_t> prof_free;
	double prof_ms[SMD_NPHASES] = {0};
	long long prof_count[SMD_NPHASES] = {0};
};
