// host_tables.hpp -- host-side construction of the per-thread twiddle tables consumed by
// pbs_team.cuh (TwTables).  Pure C++ (no CUDA), shared by the product library and tests/emu.
#pragma once
#include <stdint.h>

#include <vector>

#include "tfhe_core.cuh"

namespace tfhe {

struct HostTw {
    // index [prime]; element = (w, shoup(w)) pairs flattened as {w, ws, w, ws, ...}
    std::vector<uint32_t> fwdB[2], fwdC[2], invB[2], invC[2];
};

inline void build_tw_tables(int logn, int loge, HostTw &out) {
    const int qb = logn - 2 * loge;
    const int nb = (1 << qb) - 1, nc = (1 << loge) - 1;
    const int T = 1 << (logn - loge);
    for (int pr = 0; pr < 2; pr++) {
        const uint32_t q = prime_c(pr);
        auto push = [&](std::vector<uint32_t> &v, uint32_t w) {
            v.push_back(w);
            v.push_back(shoup_c(w, q));
        };
        out.fwdB[pr].clear(); out.invB[pr].clear(); out.fwdC[pr].clear(); out.invC[pr].clear();
        for (uint32_t hA = 0; hA < (1u << loge); hA++)
            for (int u = 0; u < qb; u++)
                for (uint32_t m = 0; m < (1u << u); m++) {
                    const uint32_t w = fwd_tw_c(pr, logn, loge + u, (hA << u) | m);
                    push(out.fwdB[pr], w);
                    push(out.invB[pr], invmod_c(w, q));
                }
        for (uint32_t t = 0; t < (uint32_t)T; t++)
            for (int u = 0; u < loge; u++)
                for (uint32_t m = 0; m < (1u << u); m++) {
                    const uint32_t w = fwd_tw_c(pr, logn, loge + qb + u, (t << u) | m);
                    push(out.fwdC[pr], w);
                    push(out.invC[pr], invmod_c(w, q));
                }
        (void)nb; (void)nc;
    }
}

}  // namespace tfhe
