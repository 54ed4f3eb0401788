// tool_common.hpp -- shared plumbing of the two host shells: plotfile metadata -> pa_level_desc, pinned level
// buffers, upload / download, error handling (errors end the process like amrex::Abort).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/pele_stencil_b200.h"
#include "parmparse.hpp"
#include "plotfile.hpp"

inline void pa_abort(const std::string& msg) {
    std::cerr << "amrex::Abort::0::" << msg << " !!!" << std::endl;      // same banner the reference prints
    std::exit(1);
}
inline void check(int rc, const char* what) {
    if (rc != PA_OK) pa_abort(std::string(what) + ": " + pa_last_error());
}
inline std::string file_root(const std::string& infile) {          // getFileRoot (grad.cpp:26-31)
    std::string s = infile;
    while (!s.empty() && s.back() == '/') s.pop_back();
    auto p = s.find_last_of('/');
    return p == std::string::npos ? s : s.substr(p + 1);
}

struct HierInput {
    std::vector<pa_level_desc> lv;
    std::vector<std::vector<int>> boxes;
};
inline void make_level_descs(const pltio::Header& h, int nlev, HierInput& out) {
    out.lv.resize(nlev);
    out.boxes.resize(nlev);
    for (int l = 0; l < nlev; ++l) {
        const auto& L = h.levels[l];
        for (int d = 0; d < 3; ++d) {
            out.lv[l].domain_lo[d] = L.domain.lo[d];
            out.lv[l].domain_hi[d] = L.domain.hi[d];
            // Geometry::CellSize = (prob_hi - prob_lo) / N  (AMReX_Geometry.cpp:520), not the rounded header value
            out.lv[l].dx[d] = (h.prob_hi[d] - h.prob_lo[d]) / (double)(L.domain.hi[d] - L.domain.lo[d] + 1);
        }
        for (auto& b : L.boxes) {
            for (int d = 0; d < 3; ++d) out.boxes[l].push_back(b.lo[d]);
            for (int d = 0; d < 3; ++d) out.boxes[l].push_back(b.hi[d]);
        }
        out.lv[l].nboxes = (int)L.boxes.size();
        out.lv[l].boxes = out.boxes[l].data();
        out.lv[l].owner = nullptr;
    }
}

// host buffer holding [comp][level cells].  Pageable by default: measured on the B200 box (profiles/r02_tool_walltime.txt),
// cudaHostAlloc of the tools' buffers costs 0.4 s per GB and as much again to release at exit -- more than the pinned DMA saves on
// data that crosses PCIe once (2 GB pinned for a 50 M-cell grad run: 0.9 s to pin, against 0.2 s of slower copies).  PA_TOOL_PIN=1
// pins them (worth it when the same buffers are reused many times).
struct PinnedLevel {
    double* p = nullptr;
    long long ncells = 0;
    int ncomp = 0;
    bool pinned = false;
    void alloc(long long cells, int comps) {
        ncells = cells; ncomp = comps;
        const char* e = std::getenv("PA_TOOL_PIN");
        pinned = e && e[0] == '1';
        const size_t bytes = (size_t)cells * comps * 8;
        if (pinned) {
            void* q = nullptr;
            check(pa_host_alloc(&q, bytes), "pa_host_alloc");
            p = (double*)q;
        } else {
            void* q = nullptr;
            if (posix_memalign(&q, 4096, bytes ? bytes : 8) != 0) pa_abort("out of host memory");
            p = (double*)q;
        }
    }
    double* comp(int c) { return p + (long long)c * ncells; }
    ~PinnedLevel() { if (p) { if (pinned) pa_host_free(p); else std::free(p); } }
};
