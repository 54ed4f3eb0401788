#include "plotfile.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace pltio {

static std::vector<int> ints_of(const std::string& s) {
    std::vector<int> v;
    size_t i = 0;
    while (i < s.size()) {
        if (isdigit((unsigned char)s[i]) || (s[i] == '-' && i + 1 < s.size() && isdigit((unsigned char)s[i + 1]))) {
            size_t j = i + 1;
            while (j < s.size() && isdigit((unsigned char)s[j])) ++j;
            v.push_back(std::atoi(s.substr(i, j - i).c_str()));
            i = j;
        } else ++i;
    }
    return v;
}

static std::vector<std::string> lines_of(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<std::string> L;
    std::string s;
    while (std::getline(f, s)) L.push_back(s);
    return L;
}

static void read_cell_h(const std::string& dir, LevelMeta& lv) {
    auto L = lines_of(dir + "/" + lv.cell_path + "_H");
    if (L.size() < 6) throw std::runtime_error("short Cell_H");
    int vers = std::atoi(L[0].c_str());
    lv.ncomp_on_disk = std::atoi(L[2].c_str());
    auto nb = ints_of(L[4]);
    int nboxes = nb.at(0);
    lv.boxes.resize(nboxes);
    for (int b = 0; b < nboxes; ++b) {
        auto v = ints_of(L.at(5 + b));
        for (int d = 0; d < 3; ++d) { lv.boxes[b].lo[d] = v.at(d); lv.boxes[b].hi[d] = v.at(3 + d); }
    }
    size_t p = 5 + nboxes + 1;
    int nf = std::atoi(L.at(p).c_str());
    ++p;
    lv.fab_file.resize(nf); lv.fab_offset.resize(nf);
    for (int i = 0; i < nf; ++i) {
        std::istringstream is(L.at(p + i));
        std::string tag;
        is >> tag >> lv.fab_file[i] >> lv.fab_offset[i];
    }
    p += nf;
    if (vers == 1) {
        for (int t = 0; t < 2; ++t) {
            while (p < L.size() && L[p].find_first_not_of(" \t\r") == std::string::npos) ++p;
            if (p >= L.size()) break;
            int n = 0, m = 0;
            if (std::sscanf(L[p].c_str(), "%d,%d", &n, &m) != 2) break;
            ++p;
            auto& tab = t == 0 ? lv.fab_min : lv.fab_max;
            tab.assign(n, std::vector<double>(m));
            for (int i = 0; i < n; ++i) {
                std::string s = L.at(p + i);
                std::replace(s.begin(), s.end(), ',', ' ');
                std::istringstream is(s);
                for (int j = 0; j < m; ++j) is >> tab[i][j];
            }
            p += n;
        }
    }
}

Header read_header(const std::string& dir) {
    auto L = lines_of(dir + "/Header");
    Header h;
    size_t p = 1;
    int nvar = std::atoi(L.at(p++).c_str());
    for (int i = 0; i < nvar; ++i) {
        std::string s = L.at(p++);
        while (!s.empty() && (s.back() == ' ' || s.back() == '\r')) s.pop_back();
        h.names.push_back(s);
    }
    int dim = std::atoi(L.at(p++).c_str());
    if (dim != 3) throw std::runtime_error("only 3-D plotfiles are supported");
    h.time = std::atof(L.at(p++).c_str());
    h.finest_level = std::atoi(L.at(p++).c_str());
    { std::istringstream is(L.at(p++)); for (int d = 0; d < 3; ++d) is >> h.prob_lo[d]; }
    { std::istringstream is(L.at(p++)); for (int d = 0; d < 3; ++d) is >> h.prob_hi[d]; }
    h.ref_ratio = ints_of(L.at(p++));
    auto dom = ints_of(L.at(p++));
    ++p;  // level steps
    h.levels.resize(h.finest_level + 1);
    for (int l = 0; l <= h.finest_level; ++l) {
        for (int d = 0; d < 3; ++d) { h.levels[l].domain.lo[d] = dom.at(9 * l + d); h.levels[l].domain.hi[d] = dom.at(9 * l + 3 + d); }
        std::istringstream is(L.at(p++));
        for (int d = 0; d < 3; ++d) is >> h.levels[l].dx[d];
    }
    h.coord = std::atoi(L.at(p++).c_str());
    ++p;  // "0"
    for (int l = 0; l <= h.finest_level; ++l) {
        std::istringstream is(L.at(p++));
        int lev, ng; double t;
        is >> lev >> ng >> t;
        ++p;  // step
        p += 3 * (size_t)ng;
        std::string path = L.at(p++);
        while (!path.empty() && (path.back() == ' ' || path.back() == '\r')) path.pop_back();
        h.levels[l].cell_path = path;
        read_cell_h(dir, h.levels[l]);
    }
    return h;
}

// One FAB on disk (AMReX_FArrayBox.cpp: FABio_binary / AMReX_FabConv.cpp): an ASCII line
//     FAB ((nfmt, (bits expbits mantbits ...)),(nord, (byte order))) ((lo) (hi) (type)) ncomp
// followed by ncomp * box.npts() reals in that RealDescriptor's format.  The descriptor may be IEEE double or IEEE single
// (fab.format = NATIVE_32 / IEEE32), little- or big-endian, and the FAB's own box may be the valid box grown by the
// MultiFab's ghost width (Cell_H ngrow > 0).  What AmrData does through RealDescriptor::convertToNativeFormat is done here
// for these four IEEE layouts (single -> double widening is exact); anything else is an error -- never a silent misread.
struct FabOnDisk { int bytes = 8; bool big_endian = false; BoxI box; int ncomp = 0; };

static FabOnDisk parse_fab_header(FILE* f, const std::string& where) {
    std::string line;
    int c;
    while ((c = std::fgetc(f)) != EOF && c != '\n') line.push_back((char)c);
    if (line.compare(0, 3, "FAB") != 0) throw std::runtime_error("not a FAB header in " + where + ": " + line.substr(0, 60));
    // the descriptor ends where the box begins: the first "((" after the leading one
    const size_t box_at = line.find("((", 5);
    if (box_at == std::string::npos) throw std::runtime_error("malformed FAB header in " + where);
    const std::vector<int> d = ints_of(line.substr(0, box_at)), b = ints_of(line.substr(box_at));
    // d = nfmt, fmt[nfmt], nord, ord[nord];  b = lo[3], hi[3], type[3], ncomp
    if (d.size() < 2 || (int)d.size() < 2 + d[0] || (int)d.size() != 2 + d[0] + d[1 + d[0]] || b.size() != 10)
        throw std::runtime_error("malformed FAB header in " + where + ": " + line);
    const std::vector<int> fmt(d.begin() + 1, d.begin() + 1 + d[0]), ord(d.begin() + 2 + d[0], d.end());
    FabOnDisk F;
    const std::vector<int> f64{64, 11, 52, 0, 1, 12, 0, 1023}, f32{32, 8, 23, 0, 1, 9, 0, 127};
    if (fmt == f64) F.bytes = 8; else if (fmt == f32) F.bytes = 4;
    else throw std::runtime_error("unsupported real format in " + where + " (only IEEE double / single): " + line.substr(0, box_at));
    if ((int)ord.size() != F.bytes) throw std::runtime_error("malformed byte order in " + where);
    bool le = true, be = true;
    for (int i = 0; i < F.bytes; ++i) { le &= ord[i] == F.bytes - i; be &= ord[i] == i + 1; }
    if (!le && !be) throw std::runtime_error("unsupported byte order in " + where + ": " + line.substr(0, box_at));
    F.big_endian = be && !le;
    for (int k = 0; k < 3; ++k) { F.box.lo[k] = b[k]; F.box.hi[k] = b[3 + k]; if (b[6 + k] != 0) throw std::runtime_error("not a cell-centred FAB in " + where); }
    F.ncomp = b[9];
    return F;
}

// component `comp` of the VALID box `vb` of the FAB at `offset` -> dst (FArrayBox order, doubles)
static void read_fab_comp(FILE* f, long offset, const BoxI& vb, int comp, double* dst, const std::string& where) {
    std::fseek(f, offset, SEEK_SET);
    const FabOnDisk F = parse_fab_header(f, where);
    if (comp < 0 || comp >= F.ncomp) throw std::runtime_error("component out of range in " + where);
    for (int d = 0; d < 3; ++d)
        if (F.box.lo[d] > vb.lo[d] || F.box.hi[d] < vb.hi[d]) throw std::runtime_error("FAB box does not contain the valid box in " + where);
    const long long nf = F.box.npts(), n = vb.npts();
    const bool same = nf == n;
    std::fseek(f, (long)((long long)F.bytes * nf * comp), SEEK_CUR);
    if (same && F.bytes == 8 && !F.big_endian) {                      // the common case: straight into the destination
        if ((long long)std::fread(dst, 8, (size_t)n, f) != n) throw std::runtime_error("short read in " + where);
        return;
    }
    std::vector<unsigned char> raw((size_t)nf * F.bytes);
    if ((long long)std::fread(raw.data(), F.bytes, (size_t)nf, f) != nf) throw std::runtime_error("short read in " + where);
    auto value = [&](long long i) -> double {
        unsigned char t[8];
        const unsigned char* p = raw.data() + (size_t)i * F.bytes;
        for (int k = 0; k < F.bytes; ++k) t[k] = F.big_endian ? p[F.bytes - 1 - k] : p[k];
        if (F.bytes == 8) { double v; std::memcpy(&v, t, 8); return v; }
        float v; std::memcpy(&v, t, 4); return (double)v;
    };
    const long long fx = F.box.hi[0] - F.box.lo[0] + 1, fy = F.box.hi[1] - F.box.lo[1] + 1;
    const int nx = vb.hi[0] - vb.lo[0] + 1, ny = vb.hi[1] - vb.lo[1] + 1, nz = vb.hi[2] - vb.lo[2] + 1;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i)
                *dst++ = value(((long long)(k + vb.lo[2] - F.box.lo[2]) * fy + (j + vb.lo[1] - F.box.lo[1])) * fx + (i + vb.lo[0] - F.box.lo[0]));
}

void read_boxes_comp(const std::string& dir, const Header& h, int lev, int comp, const std::vector<int>& box_ids, double* dst) {
    const LevelMeta& lv = h.levels.at(lev);
    std::string ldir = dir + "/" + lv.cell_path.substr(0, lv.cell_path.find_last_of('/'));
    std::string open_name;
    FILE* f = nullptr;
    try {
        for (int b : box_ids) {
            if (lv.fab_file.at(b) != open_name) {
                if (f) std::fclose(f);
                f = std::fopen((ldir + "/" + lv.fab_file[b]).c_str(), "rb");
                if (!f) throw std::runtime_error("cannot open " + ldir + "/" + lv.fab_file[b]);
                open_name = lv.fab_file[b];
            }
            read_fab_comp(f, lv.fab_offset[b], lv.boxes[b], comp, dst, ldir + "/" + open_name);
            dst += lv.boxes[b].npts();
        }
    } catch (...) {
        if (f) std::fclose(f);
        throw;
    }
    if (f) std::fclose(f);
}

void read_level_comp(const std::string& dir, const Header& h, int lev, int comp, double* dst) {
    std::vector<int> all(h.levels.at(lev).boxes.size());
    for (size_t b = 0; b < all.size(); ++b) all[b] = (int)b;
    read_boxes_comp(dir, h, lev, comp, all, dst);
}

static std::string g17(double x) { char b[64]; std::snprintf(b, sizeof b, "%.17g", x); return b; }
static std::string boxstr(const BoxI& b) {
    char s[160];
    std::snprintf(s, sizeof s, "((%d,%d,%d) (%d,%d,%d) (0,0,0))", b.lo[0], b.lo[1], b.lo[2], b.hi[0], b.hi[1], b.hi[2]);
    return s;
}

void create_plotfile_dirs(const std::string& dir, int nlev) {
    struct stat sb;
    if (::stat(dir.c_str(), &sb) == 0) {
        auto t = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
        std::string old = dir + ".old." + std::to_string(t % 10000000000LL);
        if (std::rename(dir.c_str(), old.c_str()) != 0) throw std::runtime_error("cannot rename existing " + dir);
    }
    if (::mkdir(dir.c_str(), 0755) != 0) throw std::runtime_error("cannot create " + dir);
    for (int l = 0; l < nlev; ++l) {
        std::string ldir = dir + "/Level_" + std::to_string(l);
        if (::mkdir(ldir.c_str(), 0755) != 0) throw std::runtime_error("cannot create " + ldir);
    }
}

std::vector<FabRecord> write_fab_file(const std::string& dir, const Header& meta, int lev, const std::string& file,
                                      const std::vector<int>& box_ids, const std::vector<const double*>& data) {
    const LevelMeta& lv = meta.levels.at(lev);
    const int ncomp = (int)data.size();
    std::vector<FabRecord> recs;
    if (box_ids.empty()) return recs;
    FILE* f = std::fopen((dir + "/Level_" + std::to_string(lev) + "/" + file).c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + file);
    long long cell0 = 0;
    for (int b : box_ids) {
        FabRecord r;
        r.lev = lev; r.box = b; r.file = file; r.offset = std::ftell(f);
        r.mn.resize(ncomp); r.mx.resize(ncomp);
        std::fprintf(f, "FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))%s %d\n", boxstr(lv.boxes.at(b)).c_str(), ncomp);
        const long long n = lv.boxes[b].npts();
        for (int c = 0; c < ncomp; ++c) {
            const double* p = data[c] + cell0;
            std::fwrite(p, 8, (size_t)n, f);
            double a = std::numeric_limits<double>::max(), z = std::numeric_limits<double>::lowest();
            for (long long q = 0; q < n; ++q) { a = std::min(a, p[q]); z = std::max(z, p[q]); }
            r.mn[c] = a; r.mx[c] = z;
        }
        cell0 += n;
        recs.push_back(std::move(r));
    }
    std::fclose(f);
    return recs;
}

void write_metadata(const std::string& dir, const Header& meta, const std::vector<std::string>& names, int nlev,
                    const std::vector<FabRecord>& records, const std::vector<int>& ref_ratio_line) {
    const int ncomp = (int)names.size();
    {
        std::ofstream h(dir + "/Header");
        h << "HyperCLaw-V1.1\n" << ncomp << '\n';
        for (auto& n : names) h << n << '\n';
        h << "3\n" << g17(meta.time) << '\n' << nlev - 1 << '\n';
        for (int d = 0; d < 3; ++d) h << g17(meta.prob_lo[d]) << ' ';
        h << '\n';
        for (int d = 0; d < 3; ++d) h << g17(meta.prob_hi[d]) << ' ';
        h << '\n';
        for (int l = 0; l < nlev - 1; ++l) h << ref_ratio_line.at(l) << ' ';
        h << '\n';
        for (int l = 0; l < nlev; ++l) h << boxstr(meta.levels[l].domain) << ' ';
        h << '\n';
        for (int l = 0; l < nlev; ++l) h << "0 ";
        h << '\n';
        for (int l = 0; l < nlev; ++l) { for (int d = 0; d < 3; ++d) h << g17(meta.levels[l].dx[d]) << ' '; h << '\n'; }
        h << meta.coord << "\n0\n";
        for (int l = 0; l < nlev; ++l) {
            const LevelMeta& lv = meta.levels[l];
            h << l << ' ' << lv.boxes.size() << ' ' << g17(meta.time) << "\n0\n";
            for (auto& b : lv.boxes)
                for (int d = 0; d < 3; ++d)
                    h << g17(meta.prob_lo[d] + lv.dx[d] * (b.lo[d] - lv.domain.lo[d])) << ' '
                      << g17(meta.prob_lo[d] + lv.dx[d] * (b.hi[d] - lv.domain.lo[d] + 1)) << '\n';
            h << "Level_" << l << "/Cell\n";
        }
    }
    for (int l = 0; l < nlev; ++l) {
        const LevelMeta& lv = meta.levels[l];
        std::vector<const FabRecord*> of(lv.boxes.size(), nullptr);
        for (auto& r : records) if (r.lev == l) of.at(r.box) = &r;
        for (auto* r : of) if (!r) throw std::runtime_error("write_metadata: a box of level " + std::to_string(l) + " was not written");
        std::ofstream c(dir + "/Level_" + std::to_string(l) + "/Cell_H");
        c << "1\n1\n" << ncomp << "\n0\n(" << lv.boxes.size() << " 0\n";
        for (auto& b : lv.boxes) c << boxstr(b) << '\n';
        c << ")\n" << lv.boxes.size() << '\n';
        for (auto* r : of) c << "FabOnDisk: " << r->file << ' ' << r->offset << '\n';
        c << '\n';
        for (int which = 0; which < 2; ++which) {
            c << lv.boxes.size() << ',' << ncomp << '\n';
            for (auto* r : of) {
                for (double v : (which ? r->mx : r->mn)) { char s[64]; std::snprintf(s, sizeof s, "%.17e,", v); c << s; }
                c << '\n';
            }
            c << '\n';
        }
    }
}

void write_plotfile(const std::string& dir, const Header& meta, const std::vector<std::string>& names,
                    const std::vector<std::vector<const double*>>& data, const std::vector<int>& ref_ratio_line) {
    const int nlev = (int)data.size();
    create_plotfile_dirs(dir, nlev);
    // Every level has its own Cell_D file and the write is a CPU-bound copy plus a min / max scan: one thread per box range
    // (up to four files per level, the layout VisMF produces with several writers), all levels at once.  Measured on the
    // B200 box: 0.8-0.9 s for the 2 GB of a 50 M-cell grad output from one thread.
    struct Job { int lev; std::string file; std::vector<int> boxes; std::vector<FabRecord> recs; std::string err; };
    std::vector<Job> jobs;
    for (int l = 0; l < nlev; ++l) {
        const int nb = (int)meta.levels[l].boxes.size();
        const long long cells = meta.levels[l].ncells();
        int nfiles = std::max(1, std::min(std::min(4, nb), (int)(cells * (long long)data[l].size() / (32LL << 20)) + 1));
        if (const char* e = std::getenv("PA_PLT_NFILES")) nfiles = std::max(1, std::min(nb, std::atoi(e)));   // tests: force the multi-file layout
        for (int f = 0; f < nfiles; ++f) {
            Job j;
            j.lev = l;
            char nm[32];
            std::snprintf(nm, sizeof nm, "Cell_D_%05d", f);
            j.file = nm;
            for (int b = (int)((long long)nb * f / nfiles); b < (int)((long long)nb * (f + 1) / nfiles); ++b) j.boxes.push_back(b);
            jobs.push_back(std::move(j));
        }
    }
    // data[l][c] is the level's boxes concatenated in box order: a job's boxes start at the prefix of the boxes before them
    std::vector<std::thread> th;
    for (Job& j : jobs)
        th.emplace_back([&dir, &meta, &data, &j]() {
            try {
                long long cell0 = 0;
                for (int b = 0; b < j.boxes.front(); ++b) cell0 += meta.levels[j.lev].boxes[b].npts();
                std::vector<const double*> shifted;
                // write_fab_file indexes data[c] from the first box it is given
                for (const double* p : data[j.lev]) shifted.push_back(p + cell0);
                j.recs = write_fab_file(dir, meta, j.lev, j.file, j.boxes, shifted);
            } catch (std::exception& e) { j.err = e.what(); }
        });
    for (auto& t : th) t.join();
    std::vector<FabRecord> records;
    for (Job& j : jobs) {
        if (!j.err.empty()) throw std::runtime_error(j.err);
        records.insert(records.end(), j.recs.begin(), j.recs.end());
    }
    write_metadata(dir, meta, names, nlev, records, ref_ratio_line);
}

}  // namespace pltio
