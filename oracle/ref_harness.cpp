// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// ref_harness: drives the UNMODIFIED SoftMold reference headers (taken where they lie,
// -I$(REF) at compile time; no reference source is copied into this repo) and dumps the
// per-term quantities of the MD hot path at full FP64 precision, so that
//   (a) the plain-C restatement in oracle/oracle.c can be pinned against the real thing,
//   (b) small golden fixtures can be generated for tests/golden/ (see oracle/make_golden.py).
//
// Mirrors the call sequence of MD.cpp:152-262 (CellOpt build + computeForce, molecule switch)
// and MD.cpp:589-678 (dPotential terms of the MC box move).
//
// Build: see oracle/Makefile (g++ -O3 -fopenmp -std=c++11, the reference's own flags,
// no -march=native / -ffast-math so no FMA contraction happens).
//
// Output container ("SMDG1"): repeated records
//   char name[32]; char dtype ('d' = float64, 'i' = int32); int64 count; raw little-endian data

#define LOW_DENSITY          // as MD.cpp:34 -- must precede the includes
#define CELL_SIZE_FAILURE    // as MD.cpp:13
// CellOpt::cells / hashIndices are private: the harness is compiled with -fno-access-control (read-only
// peek; the headers stay unmodified)
#include "include/MD.h"
#include "include/system.h"

#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

static FILE *out = NULL;

static void put(const char *name, char dtype, int64_t count, const void *data)
{
	char nm[32];
	memset(nm, 0, sizeof nm);
	strncpy(nm, name, 31);
	fwrite(nm, 1, 32, out);
	fwrite(&dtype, 1, 1, out);
	fwrite(&count, sizeof count, 1, out);
	fwrite(data, dtype == 'd' ? 8 : 4, (size_t)count, out);
}

static void putd(const char *name, double v) { put(name, 'd', 1, &v); }

static void put_acc(const char *name, Blob<double> &S)
{
	int n = S.readNParticles();
	std::vector<double> buf(3 * (size_t)n);
	threeVector<double> *a = S.getAccelerations();
	for (int i = 0; i < n; i++) { buf[3*i] = a[i].x; buf[3*i+1] = a[i].y; buf[3*i+2] = a[i].z; }
	put(name, 'd', 3 * (int64_t)n, buf.data());
}

static void zero_acc(Blob<double> &S)
{
	threeVector<double> *a = S.getAccelerations();
	for (int i = 0; i < S.readNParticles(); i++) { a[i].x = 0; a[i].y = 0; a[i].z = 0; }
}

typedef CellOpt<double, Potential<double>, Force<double> > PairEngine;

// dump: all per-term quantities on the configuration stored in <name>.mpd
static int cmd_dump(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: ref_harness dump name out.bin [sx sy sz]\n"); return 2; }
	const char *name = argv[2];
	threeVector<double> scale;
	scale.x = 1.0005; scale.y = 1.0005; scale.z = 1.0 / (1.0005 * 1.0005);
	if (argc >= 7) { scale.x = atof(argv[4]); scale.y = atof(argv[5]); scale.z = atof(argv[6]); }
	// 8th argument "substrate": the molecule switch of MDsubstrate.cpp:213-262 instead of MD.cpp:414-478 (OFFSET_BOUNDARY,
	// RIGIDBEND, PULLBEAD act; FLOATING_BASE, ZTORQUE, ZPOWERPOTENTIAL, NANOCORE, BALL are ignored)
	const bool substrate = argc >= 8 && !strcmp(argv[7], "substrate");

	Blob<double> S;
	Script<double, Blob<double> > io(name, std::ios::in, &S);
	io.read();
	io.close();
	out = fopen(argv[3], "wb");
	if (!out) { perror(argv[3]); return 1; }
	fwrite("SMDG1\0\0\0", 1, 8, out);

	int n = S.readNParticles();
	int nT = S.readNTypes();
	threeVector<double> size = S.readSize();
	double box[3] = { size.x, size.y, size.z };
	int hdr[3] = { n, nT, S.readNMolecules() };
	put("n_nT_nMol", 'i', 3, hdr);
	put("box", 'd', 3, box);
	putd("cutoff", S.readCutoff());
	double sc[3] = { scale.x, scale.y, scale.z };
	put("scale", 'd', 3, sc);

	{
		std::vector<double> xyz(3 * (size_t)n), vel(3 * (size_t)n);
		std::vector<int> type(n);
		position<double> *p = S.getPositions();
		threeVector<double> *v = S.getVelocities();
		for (int i = 0; i < n; i++) {
			xyz[3*i] = p[i].x; xyz[3*i+1] = p[i].y; xyz[3*i+2] = p[i].z; type[i] = p[i].type;
			vel[3*i] = v[i].x; vel[3*i+1] = v[i].y; vel[3*i+2] = v[i].z;
		}
		put("xyz", 'd', 3 * (int64_t)n, xyz.data());
		put("vel", 'd', 3 * (int64_t)n, vel.data());
		put("type", 'i', n, type.data());
	}
	put("fC", 'd', 6 * nT * nT, S.getTwoBodyFconst());
	put("uC", 'd', 6 * nT * nT, S.getTwoBodyUconst());

	PairEngine pair(S.getPositions(), S.getAccelerations(), S.getTwoBodyFconst(), S.getTwoBodyUconst(),
	                n, nT, S.readSize(), S.readPeriodic(), S.readCutoff());
	zero_acc(S);
	pair.build();
	{
		// cell id of every particle and the linked-list successor (cellOpt.h:530-585)
		std::vector<int> key(n), next(n);
		for (int i = 0; i < n; i++) { key[i] = pair.hashIndices[i].key; next[i] = pair.hashIndices[i].value; }
		put("cell_id", 'i', n, key.data());
		put("cell_next", 'i', n, next.data());
		int nc[4] = { pair.nCells.x, pair.nCells.y, pair.nCells.z, pair.nFullCells };
		put("nCells_nFull", 'i', 4, nc);
		put("fullCells", 'i', pair.nFullCells, pair.fullCells);
	}
	pair.computeForce();
	put_acc("a_pair", S);
	putd("U_pair", pair.computePotential());
	putd("dU_pair", pair.computeDPotential(scale));

	int nMol = S.readNMolecules();
	std::vector<int> molType(nMol);
	std::vector<double> U(nMol, 0.0), dU(nMol, 0.0);
	for (int k = 0; k < nMol; k++) {
		molType[k] = S.getMolecule()[k].readType();
		zero_acc(S);
		char nm[32];
		if (substrate) {
			switch (molType[k]) {
			case BOND:  S.doBondForce(k);  U[k] = S.doBondPotential(k);  dU[k] = S.doBondDPotential(k, scale);  break;
			case BEND:  S.doBendForce(k);  U[k] = S.doBendPotential(k);  dU[k] = S.doBendDPotential(k, scale);  break;
			case CHAIN: S.doChainForce(k); U[k] = S.doChainPotential(k); dU[k] = S.doChainDPotential(k, scale); break;
			case BEAD:  S.doBeadForce(k);  U[k] = S.doBeadPotential(k);  dU[k] = S.doBeadDPotential(k, scale);  break;
			case BOUNDARY:        S.doBoundaryForce(k);       U[k] = S.doBoundaryPotential(k); break;
			case OFFSET_BOUNDARY: S.doOffsetBoundaryForce(k); break;
			case RIGIDBEND:       S.doRigidBendForce(k);      break;
			case PULLBEAD:        S.doPullBeadForce(k);       dU[k] = S.doPullBeadDPotential(k, scale); break;
			default: break;
			}
		} else
		switch (molType[k]) {
		case BOND:  S.doBondForce(k);  U[k] = S.doBondPotential(k);  dU[k] = S.doBondDPotential(k, scale);  break;
		case BEND:  S.doBendForce(k);  U[k] = S.doBendPotential(k);  dU[k] = S.doBendDPotential(k, scale);  break;
		case CHAIN: S.doChainForce(k); U[k] = S.doChainPotential(k); dU[k] = S.doChainDPotential(k, scale); break;
		case BEAD:  S.doBeadForce(k);  U[k] = S.doBeadPotential(k);  dU[k] = S.doBeadDPotential(k, scale);  break;
		case BALL:  S.doBallForce(k);  U[k] = S.doBallPotential(k);  dU[k] = S.doBallDPotential(k, scale);  break;
		case BOUNDARY:        S.doBoundaryForce(k);     U[k] = S.doBoundaryPotential(k);     break;
		case FLOATING_BASE:   S.doFloatingBaseForce(k); U[k] = S.doFloatingBasePotential(k); break;
		case ZTORQUE:         S.doZTorqueForce(k);      U[k] = S.doZTorquePotential(k);      break;
		case ZPOWERPOTENTIAL: S.doZPowerForce(k);       U[k] = S.doZPowerPotential(k);       break;
		case NANOCORE: S.doNanoCoreForce(k); U[k] = S.doNanoCorePotential(k); dU[k] = S.doNanoCoreDPotential(k, scale); break;
		default: break;
		}
		snprintf(nm, sizeof nm, "a_mol%d", k);
		put_acc(nm, S);
	}
	if (nMol) {
		put("mol_type", 'i', nMol, molType.data());
		put("U_mol", 'd', nMol, U.data());
		put("dU_mol", 'd', nMol, dU.data());
	}
	Kinetic<double> kin(S.getVelocities(), n);
	putd("kinetic", kin.compute());
	fclose(out);
	return 0;
}

// mt: first <count> rand53() draws and randInt() draws of MTRand(seed) (MersenneTwister.h:284-341)
static int cmd_mt(int argc, char **argv)
{
	if (argc < 5) { fprintf(stderr, "usage: ref_harness mt seed count out.bin\n"); return 2; }
	unsigned long seed = strtoul(argv[2], NULL, 10);
	int count = atoi(argv[3]);
	out = fopen(argv[4], "wb");
	if (!out) { perror(argv[4]); return 1; }
	fwrite("SMDG1\0\0\0", 1, 8, out);
	MTRand a(seed), b(seed);
	std::vector<double> r53(count);
	std::vector<int> r32(count);
	for (int i = 0; i < count; i++) r53[i] = a.rand53();
	for (int i = 0; i < count; i++) r32[i] = (int)(uint32_t)b.randInt();
	put("rand53", 'd', count, r53.data());
	put("randInt", 'i', count, r32.data());
	fclose(out);
	return 0;
}

// phases: time the reference's own Verlet / Langevin / CellOpt / do*Force in MD.cpp:335-511 order plus the Metropolis
// box move of MD.cpp:589-721 every 8th step when the file has deltaLXY, no file I/O.
//   ref_harness phases name nsteps [warmup_reps timed_reps]
// Each rep is <nsteps> MD iterations.  Prints one line per timed rep: N nsteps seconds threads.
static void molecule_forces(Blob<double> &S)
{
	// the molecule switch of MD.cpp:414-478 for the molecule kinds on the hot path
	for (int k = 0; k < S.readNMolecules(); k++) {
		int t = S.getMolecule()[k].readType();
		if (t == CHAIN) S.doChainForce(k);
		else if (t == BOND) S.doBondForce(k);
		else if (t == BEND) S.doBendForce(k);
		else if (t == BEAD) S.doBeadForce(k);
	}
}

static void bead_mass(Blob<double> &S)
{
	// MD.cpp:340-355 / :480-494
	threeVector<double> *acc = S.getAccelerations();
	for (int k = 0; k < S.readNMolecules(); k++) {
		molecule<double, fourVector<int> > &m = S.getMolecule()[k];
		if (m.readType() != BEAD) continue;
		double r = m.getConstants()[BEADRADIUS];
		double mass = 4.0 * M_PI * r * r;
		for (int j = 0; j < m.readNBond(); j++) {
			int idx = m.getBonds()[j].s[0];
			acc[idx].x /= mass; acc[idx].y /= mass; acc[idx].z /= mass;
		}
	}
}

// one Metropolis box-move trial with the reference's own objects, MD.cpp:589-721 (molecule kinds of the hot path)
static bool box_move(Blob<double> &S, PairEngine &pair, Verlet<double> &integrate, MTRand &randNum)
{
	threeVector<double> size = S.readSize(), oldSize = S.readSize(), fluctuation;
	fluctuation.x = S.readDeltaLXY() * (2.0 * randNum.rand53() - 1.0);
	fluctuation.y = fluctuation.x;
	fluctuation.z = (size.x * size.y) / ((size.x + fluctuation.x) * (size.y + fluctuation.y));
	size.x += fluctuation.x; size.y += fluctuation.y; size.z *= fluctuation.z;
	threeVector<double> aSize = size;
	aSize.x /= oldSize.x; aSize.y /= oldSize.y; aSize.z /= oldSize.z;
	double dPotential = pair.computeDPotential(aSize);
	for (int k = 0; k < S.readNMolecules(); k++) {
		int t = S.getMolecule()[k].readType();
		if (t == BOND) dPotential += S.doBondDPotential(k, aSize);
		else if (t == BEND) dPotential += S.doBendDPotential(k, aSize);
		else if (t == CHAIN) dPotential += S.doChainDPotential(k, aSize);
		else if (t == BEAD) dPotential += S.doBeadDPotential(k, aSize);
	}
	if (S.readTension() != 0) dPotential += S.readTension() * ((size.x * size.y) - (oldSize.x * oldSize.y));
	double D = exp(dPotential / S.readInitialTemp());
	double randNumber = randNum.rand53();
	if (D >= randNumber || -dPotential <= 0) {
		position<double> *p = S.getPositions();
		for (int k = 0; k < S.readNParticles(); k++) { p[k].x *= aSize.x; p[k].y *= aSize.y; p[k].z *= aSize.z; }
		pair.resize(size);
		integrate.resize(size);
		S.setSize(size);
		return true;
	}
	return false;
}

static int cmd_phases(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: ref_harness phases name nsteps [warmup_reps timed_reps]\n"); return 2; }
	const char *name = argv[2];
	int nsteps = atoi(argv[3]);
	int warm = argc > 4 ? atoi(argv[4]) : 0, reps = argc > 5 ? atoi(argv[5]) : 1;
	Blob<double> S;
	Script<double, Blob<double> > io(name, std::ios::in, &S);
	io.read();
	io.close();
	int n = S.readNParticles();
	threeVector<double> *acc = S.getAccelerations();
	Verlet<double> integrate(S.getPositions(), S.getAccelerations(), S.getVelocities(), n,
	                         S.readSize(), S.readDeltaT(), S.readPeriodic(), NULL);
	Langevin<double> thermostat;
	thermostat.initialize(S.getAccelerations(), S.getVelocities(), S.getPositions(), n, S.readGamma(),
	                      S.readDeltaT(), S.readSeed());
	PairEngine pair(S.getPositions(), S.getAccelerations(), S.getTwoBodyFconst(), S.getTwoBodyUconst(),
	                n, S.readNTypes(), S.readSize(), S.readPeriodic(), S.readCutoff());
	zero_acc(S);
	pair.build();
	pair.computeForce();
	thermostat.compute(S.readInitialTemp());
	molecule_forces(S);
	// the barostat stream and cadence of MD.cpp:67,149,589 -- active when the file carries deltaLXY
	MTRand randNum(S.readSeed());
	const int resizeRate = 8;
	long step = 0, trials = 0, accepted = 0;
	for (int rep = 0; rep < warm + reps; rep++) {
		double t0 = omp_get_wtime();
		for (int i = 0; i < nsteps; i++, step++) {
			bead_mass(S);
			integrate.first();
			for (int k = 0; k < n; k++) { acc[k].x = 0; acc[k].y = 0; acc[k].z = 0; }
			thermostat.compute(S.readInitialTemp());
			pair.build();
			pair.computeForce();
			molecule_forces(S);
			bead_mass(S);
			integrate.second();
			if (step % resizeRate == 0 && step != 0 && S.readDeltaLXY() != 0) {
				trials++;
				if (box_move(S, pair, integrate, randNum)) accepted++;
			}
		}
		double t1 = omp_get_wtime();
		if (rep >= warm) { printf("%d %d %.6f %d\n", n, nsteps, t1 - t0, omp_get_max_threads()); fflush(stdout); }
	}
	fprintf(stderr, "box moves: %ld trials, %ld accepted\n", trials, accepted);
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 2) { fprintf(stderr, "usage: ref_harness dump|mt|phases ...\n"); return 2; }
	if (!strcmp(argv[1], "dump")) return cmd_dump(argc, argv);
	if (!strcmp(argv[1], "mt")) return cmd_mt(argc, argv);
	if (!strcmp(argv[1], "phases")) return cmd_phases(argc, argv);
	fprintf(stderr, "unknown command %s\n", argv[1]);
	return 2;
}
