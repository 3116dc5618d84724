// Host-side file boundary of the drop-in `MD` replacement: the `.mpd` script format.
//
// Format contract (reference root relative):
//   Script<T,Blob>::read / write     include/fileFormats/scriptFormat.h:64-95   (whitespace tokens, "<name>.mpd")
//   Blob::input  state machine       include/system.h:589-1313
//   Blob::output                     include/system.h:1315-1749                 (15 significant digits, layout below)
//   Blob::errorChecking              include/system.h:438-580
// Written from the format's behaviour, not from the reference's state machine: a straight tokenizer + a writer
// that reproduces the byte layout of Script::write (every emitted item is followed by one space, absent optional
// commands leave a lone space, rows end with "\n").
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <charconv>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

#include "../../include/softmold_b200.h"

namespace {

enum Cmd {
	GAMMA, INITIALTEMP, FINALTEMP, SEED, NTYPES, NMOLECULES, NPARTICLES, PERIODIC, CUTOFF, SIZE, INITIALTIME, FINALTIME,
	DELTAT, STOREINTERVAL, MEASUREINTERVAL, TWOBODYFCONST, TWOBODYUCONST, POSITIONS, VELOCITIES, MOLECULE, BANANA, DELTALXY,
	REMOVESOLVENT, TEMPSTEPINTERVAL, SOLVENTGAMMA, GAMMATYPE, TENSION, NCMD
};   // order = output order, system.h:292-302

const char *CMD_NAME[NCMD] = {"gamma", "initialTemp", "finalTemp", "seed", "nTypes", "nMolecules", "nParticles", "periodic",
                              "cutoff", "size", "initialTime", "finalTime", "deltaT", "storeInterval", "measureInterval",
                              "twoBodyFconst", "twoBodyUconst", "positions", "velocities", "molecule", "banana", "deltaLXY",
                              "removeSolvent", "tempStepInterval", "solventGamma", "gammaType", "tension"};

struct Molecule {
	int type = 0;
	int width = 0;               // ints per bond record
	std::vector<int> records;    // [n][width]
	std::vector<double> constants;
	int n() const { return width ? (int)records.size() / width : 0; }
};

// (constants, ints per record) per molecule type: system.h:1036-1120
bool molecule_shape(int type, int nTypes, int nBonds, int &nConst, int &width)
{
	switch (type) {
	case 6: nConst = 2; width = 2; return true;                  // BOND
	case 7: nConst = 2; width = 3; return true;                  // BEND
	case 8: nConst = 4; width = 3; return true;                  // CHAIN {start, nChains, length}
	case 9: nConst = 22 * nTypes * nTypes; width = 1; return true;   // BEAD
	case 10: nConst = 1; width = 1; return true;                 // SOLID
	case 11: nConst = 4; width = 1; return true;                 // BOUNDARY
	case 14: nConst = 4; width = 1; return true;                 // OFFSET_BOUNDARY
	case 12: nConst = 5; width = 2; return true;                 // RIGIDBEND
	case 13: nConst = 4; width = 1; return true;                 // PULLBEAD
	case 15: nConst = 6 * nTypes; width = 1; return true;        // FLOATING_BASE
	case 16: nConst = 4; width = 3; return true;                 // ZTORQUE
	case 17: nConst = 2; width = 2; return true;                 // ZPOWERPOTENTIAL
	case 18: nConst = 22 * nBonds; width = 1; return true;       // NANOCORE
	case 19: nConst = 2; width = 2; return true;                 // BALL
	default: return false;                                       // TORSION / DIHEDRAL / old types: 4 raw ints, unsupported here
	}
}

} // namespace

struct smd_mpd {
	double scalar[NCMD];
	bool present[NCMD];
	double size[3];
	double solventGamma[2];
	std::vector<double> gammaType;
	std::vector<double> fC, uC, xyz, vel;
	std::vector<int> type;
	std::vector<Molecule> mol;
	smd_mpd()
	{
		for (int i = 0; i < NCMD; i++) { scalar[i] = 0; present[i] = false; }
		scalar[PERIODIC] = 1;
		size[0] = size[1] = size[2] = 0;
		solventGamma[0] = solventGamma[1] = 0;
	}
};

namespace {

struct Fail {
	std::string msg;
};

struct Tokens {
	std::vector<std::string> t;
	size_t i = 0;
	bool more() const { return i < t.size(); }
	const std::string &next(const char *ctx)
	{
		if (i >= t.size()) throw Fail{std::string("unexpected end of file while reading ") + ctx};
		return t[i++];
	}
	double num(const char *ctx, const char *fmt)
	{
		const std::string &s = next(ctx);
		errno = 0;
		char *end = nullptr;
		double v = strtod(s.c_str(), &end);
		if (end == s.c_str()) throw Fail{std::string("Value after ") + ctx + " is not a numerical type!\nFormat:\n" + fmt};
		return v;
	}
	int integer(const char *ctx, const char *fmt)
	{
		const std::string &s = next(ctx);
		char *end = nullptr;
		long v = strtol(s.c_str(), &end, 10);
		if (end == s.c_str()) throw Fail{std::string("Value after ") + ctx + " is not a numerical type!\nFormat:\n" + fmt};
		return (int)v;
	}
};

void parse(smd_mpd &m, Tokens &tk)
{
	std::map<std::string, int> cmd;
	for (int i = 0; i < NCMD; i++) cmd[CMD_NAME[i]] = i;
	while (tk.more()) {
		std::string w = tk.next("command");
		if (w == "end") break;   // END_INPUT, scriptFormat.h:6
		auto it = cmd.find(w);
		if (it == cmd.end())
			throw Fail{w + " is not a recognized command!\nLocate command before " + w +
			           "!\nYou probably have too many positions, velocities, or molecules!\nAlso, check nMolecules and nParticles."};
		int c = it->second;
		switch (c) {
		case SEED: case NTYPES: case NMOLECULES: case NPARTICLES: {
			std::string fmt = std::string(CMD_NAME[c]) + " [integer]";
			m.scalar[c] = tk.integer(CMD_NAME[c], fmt.c_str());
			break;
		}
		case PERIODIC:
			tk.next("periodic");   // value is read but forced true, system.h:745-754
			m.scalar[c] = 1;
			break;
		case SIZE:
			for (int d = 0; d < 3; d++) m.size[d] = tk.num("size", "size [float] [float] [float]");
			break;
		case TWOBODYFCONST: case TWOBODYUCONST: {
			if (!m.present[NTYPES]) throw Fail{std::string("nTypes was not present before ") + CMD_NAME[c] + "!"};
			int nT = (int)m.scalar[NTYPES];
			std::vector<double> &v = (c == TWOBODYFCONST) ? m.fC : m.uC;
			v.resize(6 * (size_t)nT * nT);
			for (auto &x : v) x = tk.num(CMD_NAME[c], "twoBody?const\n [float]...\nCheck that nTypes is correct!");
			break;
		}
		case POSITIONS: {
			if (!m.present[NPARTICLES]) throw Fail{"nParticles was not present before positions!"};
			size_t n = (size_t)m.scalar[NPARTICLES];
			m.type.resize(n);
			m.xyz.resize(3 * n);
			for (size_t i = 0; i < n; i++) {
				m.type[i] = tk.integer("positions", "positions\n [integer] [float] [float] [float]\n...");
				for (int d = 0; d < 3; d++) m.xyz[3 * i + d] = tk.num("positions", "positions\n [integer] [float] [float] [float]\n...");
			}
			break;
		}
		case VELOCITIES: {
			if (!m.present[NPARTICLES]) throw Fail{"nParticles was not present before velocities!"};
			size_t n = (size_t)m.scalar[NPARTICLES];
			m.vel.resize(3 * n);
			for (auto &x : m.vel) x = tk.num("velocities", "velocities\n [float] [float] [float]\n...");
			break;
		}
		case MOLECULE: {
			if (!m.present[NMOLECULES]) throw Fail{"nMolecules was not present before molecules!"};
			int nMol = (int)m.scalar[NMOLECULES];
			int nT = (int)m.scalar[NTYPES];
			m.mol.clear();
			for (int k = 0; k < nMol; k++) {
				Molecule mol;
				mol.type = tk.integer("molecules", "molecule\n [type] [nBonds] [constants...] [bonds...]");
				int nBonds = tk.integer("molecules", "molecule\n [type] [nBonds] ...");
				int nConst = 0;
				if (mol.type == 9 && nT <= 0) throw Fail{"Error (Blob): nTypes undefined for BEAD molecule!"};
				if (!molecule_shape(mol.type, nT, nBonds, nConst, mol.width))
					throw Fail{"molecule type " + std::to_string(mol.type) + " (raw 4-integer record) is not supported by this reader"};
				if (nBonds < 0) throw Fail{"negative bond count in molecule"};
				mol.constants.resize(nConst);
				for (auto &x : mol.constants) x = tk.num("molecules", "molecule constants");
				mol.records.resize((size_t)nBonds * mol.width);
				for (auto &x : mol.records) x = tk.integer("molecules", "molecule bond indices");
				m.mol.push_back(mol);
			}
			break;
		}
		case BANANA:
			fprintf(stderr, "Monkeys must have messed with your script because I found a banana!\n");
			break;
		case SOLVENTGAMMA:
			for (int d = 0; d < 2; d++) m.solventGamma[d] = tk.num("solventGamma", "solventGamma [inner float] [outer float]");
			break;
		case GAMMATYPE: {
			if (!m.present[NTYPES]) throw Fail{"nTypes was not present before gammaType!"};
			m.gammaType.resize((size_t)m.scalar[NTYPES]);
			for (auto &x : m.gammaType) x = tk.num("gammaType", "gammaType [float] [float] ...");
			break;
		}
		default: {
			std::string fmt = std::string(CMD_NAME[c]) + " [float]";
			m.scalar[c] = tk.num(CMD_NAME[c], fmt.c_str());
		}
		}
		m.present[c] = true;
	}
}

// Blob::errorChecking, system.h:438-545
void validate(const smd_mpd &m)
{
	size_t n = (size_t)m.scalar[NPARTICLES];
	if (m.type.size() != n || m.vel.size() != 3 * n || m.xyz.size() != 3 * n) {
		std::ostringstream o;
		o << "Mismatch of parameters:\n\tNumber of particles assumed: " << n << "\n\tNumber of particle positions: " << m.type.size()
		  << "\n\tNumber of particle velocities: " << m.vel.size() / 3;
		throw Fail{o.str()};
	}
	for (size_t i = 0; i < n; i++)
		for (int d = 0; d < 3; d++) {
			double x = m.xyz[3 * i + d];
			if (x > m.size[d] || x < 0) {
				std::ostringstream o;
				o << "XYZ"[d] << " position of particle " << i << " is out of bounds.";
				throw Fail{o.str()};
			}
		}
	for (size_t k = 0; k < m.mol.size(); k++) {
		const Molecule &mol = m.mol[k];
		for (int j = 0; j < mol.n(); j++) {
			const int *r = &mol.records[(size_t)j * mol.width];
			std::ostringstream o;
			if (mol.type == 6 && (r[0] > (int)n || r[1] > (int)n || r[0] < 0 || r[1] < 0)) o << "BOND";
			if (mol.type == 7 && (r[0] > (int)n || r[1] > (int)n || r[2] > (int)n || r[0] < 0 || r[1] < 0 || r[2] < 0)) o << "BEND";
			if (mol.type == 8) {
				long long e = (long long)r[0] + (long long)r[1] * r[2];
				if (e > (long long)n || e < 0) o << "CHAIN";
			}
			if (!o.str().empty()) {
				o << " Molecule " << k << ", bond " << j << " is out of bounds!";
				throw Fail{o.str()};
			}
		}
	}
}

void put(std::string &out, const std::string &item)
{
	out += item;
	out += ' ';   // Script::write: file << out << ' '
}

std::string num(double v)
{
	char buf[64];
	snprintf(buf, sizeof buf, "%.15g", v);   // std::setprecision(15), default float field
	return buf;
}

// the two big blocks (positions, velocities: 6 numbers per particle, 26 MB for 240 000 particles) are formatted in
// parallel chunks with std::to_chars -- by the C++17 rule the same characters as printf("%.15g") in the C locale
inline void put_num(std::string &o, double v)
{
	char buf[64];
	if (std::isfinite(v)) {
		auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::general, 15);
		o.append(buf, r.ptr);
	} else {
		o += num(v);
	}
	o += ' ';
}

template <class F>
void parallel_text(size_t n, std::string &o, F fill)
{
	unsigned T = n < 32768 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
	if (T == 1) { fill((size_t)0, n, o); return; }
	std::vector<std::string> part(T);
	std::vector<std::thread> th;
	for (unsigned t = 0; t < T; t++)
		th.emplace_back([&, t] { part[t].reserve((n / T + 1) * 80); fill(n * t / T, n * (t + 1) / T, part[t]); });
	for (auto &x : th) x.join();
	for (auto &x : part) o += x;
}

std::string serialize(const smd_mpd &m)
{
	std::string o;
	o.reserve(64 + m.xyz.size() * 22 + m.vel.size() * 22);
	int nT = (int)m.scalar[NTYPES];
	for (int c = 0; c < NCMD; c++) {
		bool present = m.present[c];
		if (c == MOLECULE && m.mol.empty()) present = false;
		if (!present) { put(o, ""); continue; }   // absent command: output() returns "" once
		put(o, std::string(c != 0 ? "\n" : "") + CMD_NAME[c]);
		switch (c) {
		case SEED: case NTYPES: case NMOLECULES: case NPARTICLES: case PERIODIC:
			put(o, std::to_string((long long)m.scalar[c]));
			break;
		case SIZE:
			for (int d = 0; d < 3; d++) put(o, num(m.size[d]));
			break;
		case TWOBODYFCONST: case TWOBODYUCONST: {
			const std::vector<double> &v = (c == TWOBODYFCONST) ? m.fC : m.uC;
			for (size_t i = 0; i < v.size(); i++) {
				std::string s = (i == 0) ? "\n " : "";
				s += num(v[i]);
				if ((i + 1) % 6 == 0) s += "\n";
				put(o, s);
			}
			break;
		}
		case POSITIONS:
			parallel_text(m.type.size(), o, [&](size_t i0, size_t i1, std::string &t) {
				for (size_t i = i0; i < i1; i++) {
					if (i == 0) t += "\n ";
					t += std::to_string(m.type[i]);
					t += ' ';
					put_num(t, m.xyz[3 * i]);
					put_num(t, m.xyz[3 * i + 1]);
					put_num(t, m.xyz[3 * i + 2]);
					t.back() = '\n';
					t += ' ';
				}
			});
			break;
		case VELOCITIES:
			parallel_text(m.vel.size() / 3, o, [&](size_t i0, size_t i1, std::string &t) {
				for (size_t i = i0; i < i1; i++) {
					if (i == 0) t += "\n ";
					put_num(t, m.vel[3 * i]);
					put_num(t, m.vel[3 * i + 1]);
					put_num(t, m.vel[3 * i + 2]);
					t.back() = '\n';
					t += ' ';
				}
			});
			break;
		case MOLECULE:
			for (size_t k = 0; k < m.mol.size(); k++) {
				const Molecule &mol = m.mol[k];
				put(o, std::string(k == 0 ? "\n" : "") + std::to_string(mol.type) + "\t");
				put(o, std::to_string(mol.n()) + "\n");
				for (size_t j = 0; j < mol.constants.size(); j++) {
					std::string s = num(mol.constants[j]) + " ";
					if (mol.type == 9 && (j + 1) % 22 == 0) s += "\n";
					if (j + 1 == mol.constants.size()) s += "\n";
					put(o, s);
				}
				for (size_t j = 0; j < mol.records.size(); j++) {
					std::string s = std::to_string(mol.records[j]) + " ";
					if ((int)((j + 1) % mol.width) == 0) s += "\n";
					put(o, s);
				}
			}
			break;
		case BANANA:
			put(o, "\n");
			break;
		case SOLVENTGAMMA:
			for (int d = 0; d < 2; d++) put(o, num(m.solventGamma[d]));
			break;
		case GAMMATYPE:
			for (int t = 0; t < nT && t < (int)m.gammaType.size(); t++) put(o, num(m.gammaType[t]));
			break;
		default:
			put(o, num(m.scalar[c]));
		}
	}
	o += '\n';
	return o;
}

void set_err(char *err, size_t errlen, const std::string &msg)
{
	if (err && errlen) {
		strncpy(err, msg.c_str(), errlen - 1);
		err[errlen - 1] = 0;
	}
}

} // namespace

extern "C" int smd_mpd_read(const char *name, smd_mpd **out, char *err, size_t errlen)
{
	if (!name || !out) return SMD_ERR_ARG;
	std::string path = std::string(name) + ".mpd";   // scriptFormat.h:44-47
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) {
		set_err(err, errlen, "Could not open " + path + " in Script class!");
		return SMD_ERR_IO;
	}
	std::string text;
	char buf[1 << 16];
	size_t got;
	while ((got = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
	fclose(f);
	Tokens tk;
	{
		size_t i = 0, n = text.size();
		while (i < n) {
			while (i < n && isspace((unsigned char)text[i])) i++;
			size_t b = i;
			while (i < n && !isspace((unsigned char)text[i])) i++;
			if (i > b) tk.t.emplace_back(text, b, i - b);
		}
	}
	smd_mpd *m = new smd_mpd();
	try {
		parse(*m, tk);
		validate(*m);
	} catch (const Fail &e) {
		set_err(err, errlen, e.msg);
		delete m;
		return SMD_ERR_IO;
	}
	*out = m;
	return SMD_OK;
}

extern "C" int smd_mpd_write(const smd_mpd *m, const char *name, char *err, size_t errlen)
{
	if (!m || !name) return SMD_ERR_ARG;
	std::string path = std::string(name) + ".mpd";
	std::string text = serialize(*m);
	// The checkpoint is the only restart file and is rewritten in place every storeInterval (MD.cpp:373-377), here from a
	// worker thread while the run goes on: write a sibling, flush it to disk, then rename() over the old file, so that a
	// process that dies mid-write leaves the last good checkpoint behind, never a truncated one.
	std::string tmp = path + ".tmp";
	FILE *f = fopen(tmp.c_str(), "wb");
	if (!f) {
		set_err(err, errlen, "Could not open " + path + " in Script class!");
		return SMD_ERR_IO;
	}
	size_t w = fwrite(text.data(), 1, text.size(), f);
	bool ok = w == text.size() && fflush(f) == 0 && fsync(fileno(f)) == 0;
	ok = (fclose(f) == 0) && ok;
	if (!ok || rename(tmp.c_str(), path.c_str()) != 0) {
		remove(tmp.c_str());
		set_err(err, errlen, "short write to " + path);
		return SMD_ERR_IO;
	}
	return SMD_OK;
}

extern "C" void smd_mpd_free(smd_mpd *m) { delete m; }

static int find_cmd(const char *command)
{
	for (int i = 0; i < NCMD; i++)
		if (!strcmp(command, CMD_NAME[i])) return i;
	return -1;
}

extern "C" int smd_mpd_get_scalar(const smd_mpd *m, const char *command, double *value, int32_t *present)
{
	if (!m || !command) return SMD_ERR_ARG;
	int c = find_cmd(command);
	if (c < 0) return SMD_ERR_ARG;
	if (value) *value = m->scalar[c];
	if (present) *present = m->present[c] ? 1 : 0;
	return SMD_OK;
}

extern "C" int smd_mpd_set_scalar(smd_mpd *m, const char *command, double value)
{
	if (!m || !command) return SMD_ERR_ARG;
	int c = find_cmd(command);
	if (c < 0) return SMD_ERR_ARG;
	m->scalar[c] = value;
	m->present[c] = true;
	return SMD_OK;
}

extern "C" int smd_mpd_get_size(const smd_mpd *m, double size[3])
{
	if (!m || !size) return SMD_ERR_ARG;
	for (int d = 0; d < 3; d++) size[d] = m->size[d];
	return SMD_OK;
}

extern "C" int smd_mpd_set_size(smd_mpd *m, const double size[3])
{
	if (!m || !size) return SMD_ERR_ARG;
	for (int d = 0; d < 3; d++) m->size[d] = size[d];
	m->present[SIZE] = true;
	return SMD_OK;
}

extern "C" int smd_mpd_particles(smd_mpd *m, int32_t *n, double **xyz, int32_t **type, double **vel)
{
	if (!m) return SMD_ERR_ARG;
	if (n) *n = (int32_t)m->type.size();
	if (xyz) *xyz = m->xyz.data();
	if (type) *type = m->type.data();
	if (vel) *vel = m->vel.data();
	return SMD_OK;
}

extern "C" int smd_mpd_pair_tables(smd_mpd *m, int32_t *n_types, double **fC, double **uC)
{
	if (!m) return SMD_ERR_ARG;
	if (n_types) *n_types = (int32_t)m->scalar[NTYPES];
	if (fC) *fC = m->fC.empty() ? nullptr : m->fC.data();
	if (uC) *uC = m->uC.empty() ? nullptr : m->uC.data();
	return SMD_OK;
}

extern "C" int smd_mpd_n_molecules(const smd_mpd *m) { return m ? (int)m->mol.size() : 0; }

extern "C" int smd_mpd_molecule(smd_mpd *m, int32_t k, int32_t *type, int32_t *n_records, int32_t *record_width, int32_t **records,
                                int32_t *n_constants, double **constants)
{
	if (!m || k < 0 || k >= (int)m->mol.size()) return SMD_ERR_ARG;
	Molecule &mol = m->mol[k];
	if (type) *type = mol.type;
	if (n_records) *n_records = mol.n();
	if (record_width) *record_width = mol.width;
	if (records) *records = mol.records.data();
	if (n_constants) *n_constants = (int32_t)mol.constants.size();
	if (constants) *constants = mol.constants.data();
	return SMD_OK;
}

extern "C" int smd_create_from_mpd(smd_mpd *m, int32_t device, int32_t noise, int32_t track_unwrapped, smd_ctx **out)
{
	return smd_create_from_mpd_driver(m, device, noise, track_unwrapped, SMD_DRIVER_MD, out);
}

extern "C" int smd_create_from_mpd_driver(smd_mpd *m, int32_t device, int32_t noise, int32_t track_unwrapped, int32_t driver, smd_ctx **out)
{
	if (!m || !out) return SMD_ERR_ARG;
	if (driver != SMD_DRIVER_MD && driver != SMD_DRIVER_SUBSTRATE) return SMD_ERR_ARG;
	const bool sub = driver == SMD_DRIVER_SUBSTRATE;
	smd_desc d;
	memset(&d, 0, sizeof d);
	d.abi_version = SMD_ABI_VERSION;
	d.n_particles = (int32_t)m->type.size();
	d.n_types = (int32_t)m->scalar[NTYPES];
	d.device = device;
	for (int k = 0; k < 3; k++) d.box[k] = m->size[k];
	d.cutoff = m->scalar[CUTOFF];
	d.dt = m->scalar[DELTAT];
	const bool per_type = !(m->scalar[GAMMA] > 0) && !m->gammaType.empty();   // MD.cpp:129-138: gamma first, gammaType second
	d.gamma = per_type ? m->gammaType[0] : m->scalar[GAMMA];
	d.temperature = m->scalar[INITIALTEMP];
	d.seed = (uint64_t)(long long)m->scalar[SEED];
	d.noise = noise;
	d.track_unwrapped = track_unwrapped;
	d.rank = 0; d.nranks = 1;
	smd_ctx *ctx = nullptr;
	int rc = smd_create(&d, &ctx);
	if (rc) return rc;
	*out = ctx;   // handed out even on later failure so the caller can read smd_last_error and destroy
	if (m->fC.empty() || m->uC.empty()) return SMD_ERR_ARG;
	if ((rc = smd_set_pair_tables(ctx, m->fC.data(), m->uC.data()))) return rc;
	if (per_type && (rc = smd_set_gamma_type(ctx, (int32_t)m->gammaType.size(), m->gammaType.data()))) return rc;
	if ((rc = smd_set_particles(ctx, m->xyz.data(), m->type.data(), m->vel.data()))) return rc;
	for (auto &mol : m->mol) {
		const int32_t *rec = mol.records.data();
		const double *con = mol.constants.data();
		switch (mol.type) {
		case SMD_MOL_CHAIN: rc = smd_add_chain(ctx, mol.n(), rec, con); break;
		case SMD_MOL_BOND: rc = smd_add_bonds(ctx, mol.n(), rec, con); break;
		case SMD_MOL_BEND: rc = smd_add_bends(ctx, mol.n(), rec, con); break;
		case SMD_MOL_BEAD: rc = smd_add_beads(ctx, mol.n(), rec, con); break;
		case SMD_MOL_BOUNDARY: rc = smd_add_boundary(ctx, mol.n(), rec, con); break;
		// MD.cpp:414-478 evaluates these five and ignores the next three; MDsubstrate.cpp:213-262 does the opposite
		case SMD_MOL_BALL: rc = sub ? smd_add_inert(ctx, mol.type) : smd_add_ball(ctx, mol.n(), rec, con); break;
		case SMD_MOL_FLOATING_BASE: rc = sub ? smd_add_inert(ctx, mol.type) : smd_add_floating_base(ctx, mol.n(), rec, con); break;
		case SMD_MOL_ZTORQUE: rc = sub ? smd_add_inert(ctx, mol.type) : smd_add_ztorque(ctx, mol.n(), rec, con); break;
		case SMD_MOL_ZPOWERPOTENTIAL: rc = sub ? smd_add_inert(ctx, mol.type) : smd_add_zpower(ctx, mol.n(), rec, con); break;
		case SMD_MOL_NANOCORE: rc = sub ? smd_add_inert(ctx, mol.type) : smd_add_nanocore(ctx, mol.n(), rec, con); break;
		case SMD_MOL_OFFSET_BOUNDARY: rc = sub ? smd_add_offset_boundary(ctx, mol.n(), rec, con) : smd_add_inert(ctx, mol.type); break;
		case SMD_MOL_RIGIDBEND: rc = sub ? smd_add_rigidbend(ctx, mol.n(), rec, con) : smd_add_inert(ctx, mol.type); break;
		case SMD_MOL_PULLBEAD: rc = sub ? smd_add_pullbead(ctx, mol.n(), rec, con) : smd_add_inert(ctx, mol.type); break;
		// parsed and written back by the reference, but neither driver does anything with it ("Holy crap this doesn't work right now")
		case SMD_MOL_SOLID: rc = smd_add_inert(ctx, mol.type); break;
		default: rc = SMD_ERR_UNSUPPORTED;
		}
		if (rc) return rc;
	}
	return SMD_OK;
}
