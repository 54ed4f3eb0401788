// plotfile.hpp -- AMReX plotfile (HyperCLaw-V1.1 header, VisMF v1 Cell_H, native little-endian double FABs) reader and
// writer for the host shells.  Format per WriteGenericPlotfileHeader (AMReX_PlotFileUtil.cpp:73-155), VisMF::Header
// operator<< (AMReX_VisMF.cpp:274-336) and FABio_binary (AMReX_FArrayBox.cpp:905-912).
#pragma once
#include <array>
#include <string>
#include <vector>

namespace pltio {

struct BoxI { int lo[3], hi[3]; long long npts() const { return (long long)(hi[0]-lo[0]+1)*(hi[1]-lo[1]+1)*(hi[2]-lo[2]+1); } };

struct LevelMeta {
    BoxI domain;
    double dx[3];
    std::vector<BoxI> boxes;
    std::vector<std::string> fab_file;      // per box
    std::vector<long long> fab_offset;      // per box
    std::vector<std::vector<double>> fab_min, fab_max;   // [box][comp]
    std::string cell_path;                  // e.g. "Level_0/Cell"
    int ncomp_on_disk = 0;
    long long ncells() const { long long s = 0; for (auto& b : boxes) s += b.npts(); return s; }
};

struct Header {
    std::vector<std::string> names;
    double time = 0.0;
    int finest_level = 0;
    double prob_lo[3], prob_hi[3];
    std::vector<int> ref_ratio;
    int coord = 0;
    std::vector<LevelMeta> levels;
    int comp(const std::string& n) const { for (size_t i = 0; i < names.size(); ++i) if (names[i] == n) return (int)i; return -1; }
};

// throws std::runtime_error on malformed / missing files
Header read_header(const std::string& dir);
// one component of every box of a level, concatenated in box order ([nz][ny][nx] each) into dst (ncells doubles)
void read_level_comp(const std::string& dir, const Header& h, int lev, int comp, double* dst);
// the same for a subset of the level's boxes (global box ids, ascending): what one rank of a multi-GPU run reads
void read_boxes_comp(const std::string& dir, const Header& h, int lev, int comp, const std::vector<int>& box_ids, double* dst);

// Writing in pieces, so that several ranks (threads or processes) can each write the boxes they own -- the layout VisMF
// produces with one Cell_D file per rank (AMReX_VisMF.cpp:905-1005):
//   create_plotfile_dirs  once: renames an existing directory to <dir>.old.<unique> (UtilCreateCleanDirectory,
//                         AMReX_Utility.cpp:160-172), creates <dir> and <dir>/Level_l
//   write_fab_file        per rank and level: the given boxes (global ids; data[comp] = their cells concatenated) into
//                         Level_l/<file>; returns one record per box (file, offset, per-component min / max)
//   write_metadata        once: Header and every Level_l/Cell_H from the records of all ranks
struct FabRecord { int lev, box; std::string file; long long offset; std::vector<double> mn, mx; };
void create_plotfile_dirs(const std::string& dir, int nlev);
std::vector<FabRecord> write_fab_file(const std::string& dir, const Header& meta, int lev, const std::string& file,
                                      const std::vector<int>& box_ids, const std::vector<const double*>& data);
void write_metadata(const std::string& dir, const Header& meta, const std::vector<std::string>& names, int nlev,
                    const std::vector<FabRecord>& records, const std::vector<int>& ref_ratio_line);

// data[lev][comp] = concatenated box data as above.  An existing directory is renamed to <dir>.old.<unique>
// (UtilCreateCleanDirectory, AMReX_Utility.cpp:160-172).
void write_plotfile(const std::string& dir, const Header& meta, const std::vector<std::string>& names,
                    const std::vector<std::vector<const double*>>& data, const std::vector<int>& ref_ratio_line);

}  // namespace pltio
