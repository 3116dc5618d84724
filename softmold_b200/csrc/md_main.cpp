// MD_b200 -- drop-in replacement of the reference `MD <name>` executable (reference: MD.cpp, 755 lines).
//
// Reads <name>.mpd, integrates it on one B200 through the C ABI of libsoftmold_b200.so (include/softmold_b200.h --
// nothing else: no CUDA, no torch here), and writes what the reference writes, in its formats:
//   <name>.mpd                  rewritten every storeInterval with initialTime = now   (MD.cpp:373-377)
//   frames_<name>.xyz           appended every storeInterval (+ the t = 0 frame)       (MD.cpp:267-273, :379-381)
//   potential_ size_ lBond_ bend_ [beadPotential_] temp_ kinetic_ flicker_ [meanSquareDisplacement_]<name>.dat
//                               appended every measureInterval                         (dataExtraction.h:827-1693)
//   kEnergyDensity_<name>.dat   at exit                                                (dataExtraction.h:551-567)
//   resizeHist_<name>.dat       at exit when deltaLXY != 0                             (MD.cpp:731-746)
// and on stderr "time<TAB>seconds" per measure interval and the final "Resize acceptance ratio".
// The schedule (what happens at which step index, restart half kick, Metropolis box move every 8 steps with the
// MT19937 stream MTRand(seed)) follows MD.cpp:186-333 and the loop :335-729.
//
// Checkpoints, frames and observables are written by a WORKER THREAD from page-locked snapshots (smd_snapshot): one
// 240 000-particle store is 26 MB of text = 0.3 s of formatting, more than the 1 000 MD steps between two stores take
// on the device, so the device keeps stepping while the previous store is being formatted.  The files are the same,
// in the same order (one FIFO); SMD_SYNC_IO=1 does the work inline instead (A/B timing).
//
// Differences, all deliberate: the Langevin noise is the counter-based Philox stream of the library (the reference's
// depends on the OpenMP thread count); the old-style molecule kinds MD itself cannot parse (TORSION,
// DIHEDRAL, ...) are refused loudly; environment variable SMD_DEVICE picks the GPU.
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/softmold_b200.h"

namespace {

// MT19937 (Matsumoto & Nishimura 1998) with the 53-bit real of the reference's MTRand::rand53
// (include/algorithms/MersenneTwister.h:337-341): the barostat stream must be the reference's to the bit.
struct MT19937 {
	uint32_t s[624];
	int pos;
	explicit MT19937(uint32_t seed)
	{
		s[0] = seed;
		for (int i = 1; i < 624; i++) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
		pos = 624;
	}
	uint32_t u32()
	{
		if (pos >= 624) {
			for (int k = 0; k < 624; k++) {
				uint32_t y = (s[k] & 0x80000000u) | (s[(k + 1) % 624] & 0x7fffffffu);
				s[k] = s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
			}
			pos = 0;
		}
		uint32_t y = s[pos++];
		y ^= y >> 11;
		y ^= (y << 7) & 0x9d2c5680u;
		y ^= (y << 15) & 0xefc60000u;
		y ^= y >> 18;
		return y;
	}
	double rand53()
	{
		uint32_t a = u32() >> 5, b = u32() >> 6;
		return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
	}
};

struct Mol { int type, n, width; int32_t *rec; int nconst; double *c; };

struct Job {
	int kind;                 // 0: store (Script::write + xyz frame), 1: measure (dataExtraction::compute)
	bool write_mpd, diffusion;
	double time_now, temperature, box[3], terms[SMD_NTERMS], kinetic;
	int slot;
	int64_t ticket;
	// kind 1: the geometric observables, already reduced on the device (smd_observe)
	smd_observables obs;
	std::vector<double> msd_sum;
	std::vector<int64_t> msd_count;
};

struct Driver {
	std::string name;
	smd_mpd *mpd = nullptr;
	smd_ctx *ctx = nullptr;
	int n = 0;
	double *xyz = nullptr, *vel = nullptr;   // borrowed from the mpd object; written by the worker only
	int32_t *type = nullptr;
	std::vector<Mol> mols;
	bool diffusion_started = false;          // aPStart of the reference (MD.cpp:96-105, dataExtraction.h:790-803): smd_msd_start
	double temperature = 0, time_now = 0;

	// snapshots in flight: two page-locked buffer sets, one FIFO of jobs, one worker
	struct Slot { double *xyz = nullptr, *vel = nullptr, *unw = nullptr; bool busy = false; } slots[2];
	std::thread worker;
	std::mutex mu;
	std::condition_variable cv_job, cv_free;
	std::deque<Job> jobs;
	bool quit = false, sync_io = false;

	std::timed_mutex store_mu;   // held while a checkpoint is being written
	void die(const char *what, int rc)
	{
		std::cerr << what << ": " << (ctx ? smd_last_error(ctx) : smd_last_error(nullptr)) << " (code " << rc << ")\n";
		// a checkpoint the worker is writing right now describes a state from BEFORE this error: let it finish (the write
		// itself is atomic, smd_mpd_write renames a finished sibling over the old file, so this only saves the newer state)
		if (store_mu.try_lock_for(std::chrono::seconds(30))) store_mu.unlock();
		std::_Exit(1);
	}
	void ck(int rc, const char *what) { if (rc) die(what, rc); }

	void start_io()
	{
		sync_io = std::getenv("SMD_SYNC_IO") != nullptr;
		for (Slot &s : slots) {
			ck(smd_host_alloc((void **)&s.xyz, 3 * (size_t)n * sizeof(double)), "smd_host_alloc");
			ck(smd_host_alloc((void **)&s.vel, 3 * (size_t)n * sizeof(double)), "smd_host_alloc");
		}
		if (!sync_io) worker = std::thread([this] { work(); });
	}

	void stop_io()
	{
		if (worker.joinable()) {
			{ std::lock_guard<std::mutex> l(mu); quit = true; }
			cv_job.notify_all();
			worker.join();
		}
		for (Slot &s : slots) { smd_host_free(s.xyz); smd_host_free(s.vel); }
	}

	void work()
	{
		for (;;) {
			Job j;
			{
				std::unique_lock<std::mutex> l(mu);
				cv_job.wait(l, [this] { return quit || !jobs.empty(); });
				if (jobs.empty()) return;
				j = jobs.front();
				jobs.pop_front();
			}
			run(j);
			if (j.kind == 0) {
				{ std::lock_guard<std::mutex> l(mu); slots[j.slot].busy = false; }
				cv_free.notify_all();
			}
		}
	}

	// a store job: snapshot of the current state into a free buffer set + the job that consumes it; a measure job carries
	// its numbers with it.  One FIFO, so the files are appended in the reference's order.
	void submit(Job j)
	{
		if (j.kind == 1) {
			if (sync_io) { run(j); return; }
			{ std::lock_guard<std::mutex> l(mu); jobs.push_back(std::move(j)); }
			cv_job.notify_all();
			return;
		}
		{
			std::unique_lock<std::mutex> l(mu);
			cv_free.wait(l, [this] { return !slots[0].busy || !slots[1].busy; });
			j.slot = slots[0].busy ? 1 : 0;
			slots[j.slot].busy = true;
		}
		Slot &s = slots[j.slot];
		ck(smd_snapshot(ctx, s.xyz, s.vel, nullptr, &j.ticket), "smd_snapshot");
		if (sync_io) {
			run(j);
			slots[j.slot].busy = false;
			return;
		}
		{ std::lock_guard<std::mutex> l(mu); jobs.push_back(j); }
		cv_job.notify_all();
	}

	void run(const Job &j)
	{
		if (j.kind == 1) { do_measure(j); return; }
		ck(smd_snapshot_wait(ctx, j.ticket), "smd_snapshot_wait");
		do_store(j, slots[j.slot]);
	}

	// Script::write + xyzFormat::store at a store step (MD.cpp:373-381)
	void store(bool write_mpd)
	{
		Job j = {};
		j.kind = 0; j.write_mpd = write_mpd; j.time_now = time_now; j.temperature = temperature;
		// device-side errors (a particle outside the box, NaN positions) only surface at a synchronisation: a state that
		// has already raised one must not replace the last good checkpoint
		if (write_mpd) ck(smd_synchronize(ctx), "state check before the checkpoint");
		smd_get_box(ctx, j.box);
		submit(j);
	}

	void do_store(const Job &j, const Slot &s)
	{
		if (j.write_mpd) {
			std::memcpy(xyz, s.xyz, 3 * (size_t)n * sizeof(double));
			std::memcpy(vel, s.vel, 3 * (size_t)n * sizeof(double));
			smd_mpd_set_size(mpd, j.box);
			smd_mpd_set_scalar(mpd, "initialTime", j.time_now);
			smd_mpd_set_scalar(mpd, "initialTemp", j.temperature);
			char err[512];
			std::lock_guard<std::timed_mutex> hold(store_mu);
			if (smd_mpd_write(mpd, name.c_str(), err, sizeof err)) { std::cerr << err << "\n"; std::_Exit(1); }
		}
		// xyzFormat::store (xyzFormat.h:111-143): default stream precision = %g; formatted in parallel chunks
		const unsigned T = n < 32768 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
		std::vector<std::string> part(T);
		auto fill = [&](unsigned t) {
			std::string &o = part[t];
			o.reserve(((size_t)n / T + 1) * 48);
			char buf[64];
			for (size_t i = (size_t)n * t / T; i < (size_t)n * (t + 1) / T; i++) {
				o += std::to_string(type[i]);
				for (int a = 0; a < 3; a++) {
					o += '\t';
					const double v = s.xyz[3 * i + a];
					if (std::isfinite(v)) o.append(buf, std::to_chars(buf, buf + sizeof buf, v, std::chars_format::general, 6).ptr);
					else { snprintf(buf, sizeof buf, "%g", v); o += buf; }
				}
				o += '\n';
			}
		};
		if (T == 1) fill(0);
		else {
			std::vector<std::thread> th;
			for (unsigned t = 0; t < T; t++) th.emplace_back(fill, t);
			for (auto &x : th) x.join();
		}
		std::ofstream f("frames_" + name + ".xyz", std::ios::out | std::ios::app);
		f << n << '\n' << "test\n";
		for (auto &o : part) f << o;
	}

	template <class... A>
	void line(double t, const char *prefix, A... values)
	{
		std::ofstream f(prefix + name + ".dat", std::ios::app | std::ios::out);
		f << t;
		((f << '\t' << values), ...);
		f << std::endl;
	}

	// dataExtraction::compute (dataExtraction.h:827-1693, default build: no ANCHOR_DATA / FLAT_MEMBRANE / NANOPARTICLE):
	// everything is reduced on the device -- the energies by smd_potential / smd_kinetic, the bond and bend means, the extent
	// of the particles (flicker), the kinetic-energy histogram and the mean square displacements by smd_observe -- so a
	// measurement moves < 1 KB to the host; the worker only appends the lines
	void measure()
	{
		Job j = {};
		j.kind = 1; j.time_now = time_now; j.temperature = temperature; j.diffusion = diffusion_started;
		ck(smd_potential(ctx, j.terms), "smd_potential");
		ck(smd_kinetic(ctx, &j.kinetic), "smd_kinetic");
		smd_get_box(ctx, j.box);
		j.msd_sum.assign(mols.size() + 1, 0.0);
		j.msd_count.assign(mols.size() + 1, 0);
		ck(smd_observe(ctx, SMD_OBS_BONDS | SMD_OBS_EXTENT | SMD_OBS_KE_HIST | (diffusion_started ? SMD_OBS_MSD : 0u), &j.obs,
		               j.msd_sum.data(), j.msd_count.data(), (int32_t)mols.size()), "smd_observe");
		submit(std::move(j));
	}

	void do_measure(const Job &j)
	{
		const double *terms = j.terms, *s = j.box;
		const double kinetic = j.kinetic, temperature = j.temperature, time_now = j.time_now;
		double potential = 0;
		for (int t = 0; t < SMD_NTERMS; t++) potential += terms[t];
		const smd_observables &o = j.obs;
		double costhetaBend = o.cos_bend_sum, lBend[2] = {o.lbend_sum[0], o.lbend_sum[1]};
		const long long nBond = o.n_bond, nBend = o.n_bend;
		int nBeads = 0;
		for (const Mol &m : mols)
			if (m.type == SMD_MOL_BEAD || m.type == SMD_MOL_NANOCORE) nBeads++;   // dataExtraction.h:936-942, :971-978
		line(time_now, "potential_", potential);
		line(time_now, "size_", s[0], s[1], s[2]);
		line(time_now, "lBond_", o.lbond_sum / (double)nBond);
		if (nBend > 1) { costhetaBend /= nBend; lBend[0] /= nBend; lBend[1] /= nBend; }
		line(time_now, "bend_", costhetaBend, lBend[0], lBend[1]);
		if (nBeads > 0) line(time_now, "beadPotential_", terms[SMD_TERM_BEAD] + terms[SMD_TERM_NANOCORE]);
		line(time_now, "temp_", temperature);
		line(time_now, "kinetic_", kinetic);
		line(time_now, "flicker_", o.hi[0] - o.lo[0], o.hi[1] - o.lo[1], o.hi[2] - o.lo[2]);
		if (j.diffusion) {   // dataExtraction.h:1525-1663: one column per BOND / BEND / BEAD / CHAIN molecule with particles
			std::ofstream f("meanSquareDisplacement_" + name + ".dat", std::ios::app | std::ios::out);
			f << time_now;
			for (size_t k = 0; k < mols.size(); k++) {
				const int t = mols[k].type;
				if (t != SMD_MOL_BOND && t != SMD_MOL_BEND && t != SMD_MOL_BEAD && t != SMD_MOL_CHAIN) continue;
				if (j.msd_count[k] != 0) f << '\t' << (j.msd_sum[k] / (double)j.msd_count[k]);
			}
			f << '\n';
		}
	}

	void start_diffusion()
	{
		if (diffusion_started) return;
		ck(smd_msd_start(ctx), "smd_msd_start");
		diffusion_started = true;
	}

	void finish()
	{
		std::ofstream f("kEnergyDensity_" + name + ".dat", std::ios::out);
		int64_t nb = 0;
		ck(smd_ke_histogram(ctx, nullptr, 0, &nb), "smd_ke_histogram");
		std::vector<int64_t> ke_hist((size_t)nb);
		if (nb) ck(smd_ke_histogram(ctx, ke_hist.data(), nb, &nb), "smd_ke_histogram");
		long sum = 0;
		for (long c : ke_hist) sum += c;
		for (size_t i = 0; i < ke_hist.size(); i++)
			f << static_cast<float>(i) * 0.0001 << '\t' << static_cast<float>(ke_hist[i]) / static_cast<float>(sum) << std::endl;
	}
};

double scalar(smd_mpd *m, const char *cmd, bool *present = nullptr)
{
	double v = 0;
	int32_t p = 0;
	smd_mpd_get_scalar(m, cmd, &v, &p);
	if (present) *present = p != 0;
	return v;
}

}   // namespace

int main(int argc, char *argv[])
{
	if (argc != 2) {
		std::cerr << "usage: " << argv[0] << " name\n";   // MD.cpp:74-79
		return 0;
	}
	Driver D;
	D.name = argv[1];
	char err[1024];
	if (smd_mpd_read(argv[1], &D.mpd, err, sizeof err)) {
		std::cerr << err << "\n";
		return 1;
	}
	smd_mpd *M = D.mpd;
	const double gamma = scalar(M, "gamma"), dt = scalar(M, "deltaT");
	bool has_gamma_type = false;
	scalar(M, "gammaType", &has_gamma_type);
	// SMD_DRIVER: which of the reference's drivers this run stands in for -- "md" (MD.cpp, default), "anneal" (MDanneal.cpp: the
	// same loop with a box-move trial every 4th step, MDanneal.cpp:67) or "substrate" (MDsubstrate.cpp: its molecule switch,
	// :213-262, and its per-trial dAcceptFile.dat / dRejectFile.dat, :710-734)
	const char *drv = std::getenv("SMD_DRIVER");
	const std::string driver = drv ? drv : "md";
	if (driver != "md" && driver != "anneal" && driver != "substrate") {
		std::cerr << "MD_b200: SMD_DRIVER must be md, anneal or substrate\n";
		return 1;
	}
	const bool substrate = driver == "substrate";
	if (!(gamma > 0) && !has_gamma_type) {   // MD.cpp:129-143: gamma, else gammaType (smd_set_gamma_type), else give up
		(substrate ? std::cerr : std::cout) << "Error(main): No gamma available!\n";   // (MDsubstrate.cpp:143 writes it to stderr)
		return 0;
	}
	const uint32_t seed = (uint32_t)scalar(M, "seed");
	const double deltaLXY = scalar(M, "deltaLXY"), tension = scalar(M, "tension");
	const double finalTime = scalar(M, "finalTime"), initialTime = scalar(M, "initialTime");
	const double finalTemp = scalar(M, "finalTemp"), tempStepInterval = scalar(M, "tempStepInterval");
	D.temperature = scalar(M, "initialTemp");

	smd_mpd_particles(M, &D.n, &D.xyz, &D.type, &D.vel);
	for (int k = 0; k < smd_mpd_n_molecules(M); k++) {
		Mol m;
		smd_mpd_molecule(M, k, &m.type, &m.n, &m.width, &m.rec, &m.nconst, &m.c);
		D.mols.push_back(m);
	}
	const char *dev = std::getenv("SMD_DEVICE");
	int rc = smd_create_from_mpd_driver(M, dev ? std::atoi(dev) : 0, SMD_NOISE_PHILOX, 1, substrate ? SMD_DRIVER_SUBSTRATE : SMD_DRIVER_MD, &D.ctx);
	if (rc) {
		std::cerr << "MD_b200: " << smd_last_error(D.ctx) << " (code " << rc << ")\n";
		return 1;
	}
	smd_ctx *ctx = D.ctx;
	D.start_io();

	// MD.cpp:311-323: integer step indices; the 1e-7 is the reference's
	const int endInt = int(finalTime / dt + 0.0000001), startInt = int(initialTime / dt + 0.0000001);
	const int storeint = int(scalar(M, "storeInterval") / dt + 0.0000001), measureint = int(scalar(M, "measureInterval") / dt + 0.0000001);
	int tempStepInt = 0;
	double tempStep = 0;
	if (tempStepInterval > 0) {
		tempStep = tempStepInterval * (finalTemp - D.temperature) / (finalTime - initialTime);
		int div = int((finalTime - initialTime) / tempStepInterval);
		tempStepInt = div ? (endInt - startInt) / div : 0;
	}
	// MD.cpp:67 `#define resizeRate 8`; the driver variant MDanneal.cpp:67 is the same loop with `resizeRate 4` (its temperature
	// ramp is the tempStepInterval command, handled below for every run): SMD_RESIZE_RATE=4 makes this executable that variant
	int resizeRate = driver == "anneal" ? 4 : 8;
	if (const char *e = std::getenv("SMD_RESIZE_RATE")) { int v = std::atoi(e); if (v >= 1) resizeRate = v; }

	double resizeHistInterval = 0.00001;
	std::vector<double> resizeHist, rejectHist;
	if (deltaLXY != 0) {
		resizeHistInterval = deltaLXY / 101.0;
		int nIntervals = static_cast<int>(2.0 * deltaLXY / resizeHistInterval) + 1;
		resizeHist.assign(nIntervals, 0.0);
		rejectHist.assign(nIntervals, 0.0);
	}
	MT19937 randNum(seed);
	double trial = 0, accepted = 0;

	// MD.cpp:186-262: forces of the loaded configuration (pair, thermostat, molecules)
	D.time_now = initialTime;
	// (the noise of this evaluation is keyed startInt - 1: the first iteration of the loop below evaluates forces under the
	// key startInt, and two evaluations must never share their random kicks -- the reference draws fresh numbers each time)
	D.ck(smd_compute_forces(ctx, SMD_MASK_ALL, (int64_t)startInt - 1), "smd_compute_forces");
	if (initialTime == 0) {
		D.measure();
		D.store(false);
	} else {
		D.ck(smd_resume(ctx), "smd_resume");   // MD.cpp:274-308
	}

	auto store_at = [&](int i) { return storeint > 0 && i % storeint == 0 && i != startInt; };
	auto measure_at = [&](int i) { return measureint > 0 && i % measureint == 0 && i != startInt; };
	auto mc_at = [&](int i) { return i % resizeRate == 0 && i != 0 && deltaLXY != 0; };
	auto ramp_at = [&](int i) { return tempStepInterval > 0 && tempStepInt != 0 && i % tempStepInt == 0 && i < endInt; };
	auto plain = [&](int i) { return !store_at(i) && !measure_at(i) && !mc_at(i) && !ramp_at(i); };

	// one Metropolis box-move trial, MD.cpp:589-721: the two draws of MTRand randNum(seed), the trial itself (`run`), the
	// histograms of resizeHist_<name>.dat
	auto mc_trial_bookkeeping = [&](int at_step, auto run) {
		double u_fluct = randNum.rand53(), u_accept = randNum.rand53();
		double fluct = deltaLXY * (2.0 * u_fluct - 1.0);
		int32_t acc = 0;
		double dU = 0, box[3];
		run(u_fluct, u_accept, &acc, &dU, box);
		size_t bin = (size_t)((fluct + deltaLXY) / resizeHistInterval);
		if (bin < resizeHist.size()) (acc ? resizeHist : rejectHist)[bin] += 1.0;   // 0.5 for x + 0.5 for y
		if (acc) accepted++;
		trial++;
		if (substrate) {   // MDsubstrate.cpp:710-734
			std::ofstream f(acc ? "dAcceptFile.dat" : "dRejectFile.dat", std::ios::out | std::ios::app);
			f << (double)at_step * dt << '\t' << dU << std::endl;
		}
	};

	std::cerr << "starting main loop: \n";
	time_t current = time(NULL);
	const auto loop_t0 = std::chrono::steady_clock::now();
	for (int i = startInt; i <= endInt;) {
		// steps without any host-side event run back to back on the device
		int nplain = 0;
		while (i + nplain <= endInt && plain(i + nplain) && nplain < 4096) nplain++;
		// ... and when all the next eventful step does is a box-move trial, it joins them: smd_step_mc lets the pair kernel
		// of that step sum the dPotential of the proposed move along with its forces
		const int j = i + nplain;
		if (j <= endInt && mc_at(j) && !store_at(j) && !measure_at(j) && !ramp_at(j)) {
			mc_trial_bookkeeping(j, [&](double u_fluct, double u_accept, int32_t *acc, double *dU, double *box) {
				D.ck(smd_step_mc(ctx, i, nplain + 1, deltaLXY, tension, u_fluct, u_accept, acc, dU, box), "smd_step_mc");
			});
			i = j + 1;
			continue;
		}
		if (nplain > 0) {
			D.ck(smd_step(ctx, i, nplain), "smd_step");
			i += nplain;
			continue;
		}
		D.time_now = (double)i * dt;
		D.ck(smd_step_begin(ctx, i), "smd_step_begin");
		if (ramp_at(i)) {
			D.temperature += tempStep;
			smd_set_temperature(ctx, D.temperature);
		}
		if (store_at(i)) D.store(true);
		D.ck(smd_step_end(ctx, i), "smd_step_end");
		if (measure_at(i)) {
			time_t last = current;
			current = time(NULL);
			std::cerr << D.time_now << '\t' << current - last << std::endl;
			if (D.time_now > 100.0) D.start_diffusion();   // DIFFUSION_START, MD.cpp:41,538-539
			D.measure();
			current = time(NULL);
		}
		if (mc_at(i))
			mc_trial_bookkeeping(i, [&](double u_fluct, double u_accept, int32_t *acc, double *dU, double *box) {
				D.ck(smd_mc_box_move(ctx, deltaLXY, tension, u_fluct, u_accept, acc, dU, box), "smd_mc_box_move");
			});
		i++;
	}
	D.ck(smd_synchronize(ctx), "smd_synchronize");
	const auto loop_t1 = std::chrono::steady_clock::now();
	D.stop_io();   // every queued store / measurement is on disk
	if (std::getenv("SMD_TIMING")) {   // not part of the reference's output: loop seconds, then seconds until the writer had drained
		const auto loop_t2 = std::chrono::steady_clock::now();
		std::cerr << "SMD_TIMING loop " << std::chrono::duration<double>(loop_t1 - loop_t0).count() << " drained "
		          << std::chrono::duration<double>(loop_t2 - loop_t0).count() << std::endl;
	}

	if (deltaLXY != 0) {
		std::ofstream f(substrate ? std::string("resizeHist.dat") : "resizeHist_" + D.name + ".dat", std::ios::out);   // MDsubstrate.cpp:753
		for (size_t k = 0; k < resizeHist.size(); k++)
			f << (static_cast<double>(k) * resizeHistInterval) - deltaLXY << '\t' << resizeHist[k] << '\t' << rejectHist[k] << std::endl;
	}
	std::cerr << "Resize acceptance ratio: " << accepted / trial << std::endl;
	D.finish();
	smd_destroy(ctx);
	smd_mpd_free(M);
	return 0;
}
