// pbx_peer.cuh -- device side of the peer boards (PeerBoard / PeerLinks, pbx_internal.h): the
// all-reduce the ranks of a z-slab decomposition do among themselves, over NVLink peer mappings,
// inside the kernel that produced the partial sums.
#pragma once

#include "pbx_internal.h"
#include "pbx_ptx.cuh"

namespace pbx {

// Sum `count` <= PEER_VALS doubles over all ranks.  Called by EVERY thread of a CTA with at least
// L.n threads (thread r talks to rank r); `mine` (this rank's partials) and `all` live in shared
// memory and `mine` is already visible to the whole CTA.  The sums, taken in rank order (the same
// bits on every rank), are returned in result[] of thread 0.
//
// Protocol: record [seq % PEER_RING][my rank] of every rank's board receives my partials (plain
// system-scope stores, then the sequence number with release semantics); I wait for the L.n
// records of my own board to carry `seq` (acquire) and read them.  Every rank issues the same
// sequence of reductions, and nobody can complete reduction s before everybody has stored its
// record of s, i.e. has finished reading the records of s - 1: at most two consecutive ring slots
// are ever in use.
__device__ __forceinline__ void peer_exchange_sum(const PeerLinks &L, unsigned long long seq,
                                                  const double *mine, int count,
                                                  double (*all)[PEER_VALS], double *result)
{
    const int r = (int)threadIdx.x;
    const int slot = (int)(seq % PEER_RING);
    if (r < L.n) {
        PeerRec *dst = &L.board[r]->rec[slot][L.rank];
        for (int a = 0; a < count; ++a) ptx::st_relaxed_sys(&dst->v[a], mine[a]);
        ptx::st_release_sys(&dst->seq, seq);
    }
    __syncthreads();   // every store is on its way before anybody starts to wait
    if (r < L.n) {
        const PeerRec *src = &L.board[L.rank]->rec[slot][r];
        const long long t0 = ptx::spin_start();
        while (ptx::ld_acquire_sys(&src->seq) != seq) ptx::spin_pause(t0);
        for (int a = 0; a < count; ++a) all[r][a] = ptx::ld_relaxed_sys(&src->v[a]);
    }
    __syncthreads();
    if (r == 0) {
        for (int a = 0; a < count; ++a) {
            double s = 0.0;
            for (int q = 0; q < L.n; ++q) s += all[q][a];
            result[a] = s;
        }
    }
}

}  // namespace pbx
