// api_internal.hpp -- shared between host_api.cpp and tfhe_b200.cu: which kernel instantiation
// serves a parameter set.  Adding a configuration = one line here + one line in the dispatch
// tables of tfhe_b200.cu (+ one in tests/emu for the CPU cross-check).
#pragma once
#include "../../include/tfhe_b200.h"

namespace tfhe_host {

// (log2 N, k, pbs_levels, pbs_log_base) -> index into the PbsCfg instantiation list
inline int pbs_config_id(const tfhe_params &p) {
    struct Row { uint32_t logn, k, l, lb; };
    static const Row rows[] = {
        {9, 2, 6, 4},   // 0: reference defaults lib.rs:76-124 (P0 / P0t)
        {10, 1, 3, 8},  // 1: P1
        {11, 1, 3, 8},  // 2: P2
    };
    for (int i = 0; i < (int)(sizeof(rows) / sizeof(rows[0])); i++)
        if (rows[i].logn == p.glwe_poly_degree && rows[i].k == p.glwe_dimension && rows[i].l == p.pbs_levels &&
            rows[i].lb == p.pbs_log_base)
            return i;
    return -1;
}
// (ks_log_base, ks_levels) -> index into the KS decomposer instantiation list
inline int ks_config_id(const tfhe_params &p) {
    if (p.ks_log_base == 4 && p.ks_levels == 5) return 0;
    if (p.ks_log_base == 2 && p.ks_levels == 8) return 1;
    return -1;
}

}  // namespace tfhe_host
