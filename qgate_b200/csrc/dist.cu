/*
 * dist.cu — lane exchange between shards over NVLink peer memory (sm_100a).
 *
 * A state vector sharded on its g high-order (global) lanes keeps one shard of 2^(n-g)
 * amplitudes per GPU (one process each).  A non-diagonal gate on a global lane needs that
 * lane to become local: k global lanes trade places with k local "victim" lanes.  Seen from
 * rank R whose k selected global bits are s: the amplitude with local index i whose victim
 * bits are j != s belongs, after the exchange, to the rank whose selected bits are j, at the
 * local index i' = i with the victim bits replaced by s — and that rank's amplitude at i'
 * belongs here at i.  So the exchange is a set of pairwise in-place swaps between this shard
 * and 2^k - 1 peers.  This kernel does them directly through peer pointers (CUDA IPC mappings
 * of the other ranks' shards), 16 bytes per access: no staging buffer, no second copy, both
 * directions of every NVLink used at once (each rank performs half of the swaps of each pair:
 * it pulls the peer's value over the link and pushes its own).
 *
 * What it replaces in the reference: there is no exchange step — every gate on a high lane
 * streams half of every chunk over the link through fine-grained peer loads/stores inside
 * the gate kernel (MultiChunkPtr.h:30-41, ProcessorRelocator.cpp:51-73), every time.
 */
#include <cuda_runtime.h>

#include "kernels.h"

namespace qgb {

namespace {

__device__ __forceinline__ uint64_t insert_bit(uint64_t idx, int pos, uint64_t bit) {
    const uint64_t lo = idx & ((1ull << pos) - 1ull);
    return ((idx - lo) << 1) | (bit << pos) | lo;
}

/* One work item = one 16-byte unit of this shard that has to trade places.  Items are
 * numbered t = jj * 2^(n_unit_bits - k - 1) + rest, where jj enumerates the 2^k - 1 peers and
 * `rest` the unit index with the victim bits and the split bit taken out. */
__global__ void __launch_bounds__(256)
exchange_p2p_kernel(uint4 *__restrict__ local, const __grid_constant__ ExchangeParams ep) {
    const int k = ep.k, s = ep.my_sel;
    const int rest_bits = ep.n_unit_bits - k - 1;
    const uint64_t n_rest = 1ull << rest_bits;
    const uint64_t n_items = n_rest * ((1ull << k) - 1ull);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    /* positions to re-insert, ascending: the k victims and the split bit */
    int pos[QGB_MAX_EXCHANGE_LANES + 1];
    int which[QGB_MAX_EXCHANGE_LANES + 1]; /* victim number, or -1 for the split bit */
    {
        int n = 0, v = 0;
        bool split_done = false;
        while (n < k + 1) {
            if (!split_done && (v >= k || ep.split < ep.victim[v])) {
                pos[n] = ep.split;
                which[n] = -1;
                split_done = true;
            } else {
                pos[n] = ep.victim[v];
                which[n] = v;
                ++v;
            }
            ++n;
        }
    }
    /* item t -> (unit index in this shard, unit index in the peer's shard, peer) */
    auto locate = [&](uint64_t t, uint64_t &mine, uint64_t &theirs, int &j) {
        const uint64_t rest = t & (n_rest - 1ull);
        j = (int)(t >> rest_bits);
        if (j >= s) ++j; /* skip my own block */
        /* the lower-numbered rank of a pair moves the units whose split bit is 0 */
        const uint64_t split_bit = (s < j) ? 0ull : 1ull;
        mine = rest;
        theirs = rest;
#pragma unroll
        for (int n = 0; n < QGB_MAX_EXCHANGE_LANES + 1; ++n) {
            if (n < k + 1) {
                const int w = which[n];
                const uint64_t bm = (w < 0) ? split_bit : (uint64_t)((j >> w) & 1);
                const uint64_t bt = (w < 0) ? split_bit : (uint64_t)((s >> w) & 1);
                mine = insert_bit(mine, pos[n], bm);
                theirs = insert_bit(theirs, pos[n], bt);
            }
        }
    };
    /* four swaps per thread in flight: the remote loads cross NVLink (microseconds), the more of
     * them are outstanding the better the links are used */
    constexpr int UNROLL = 4;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; t + (UNROLL - 1) * stride < n_items; t += UNROLL * stride) {
        uint64_t mine[UNROLL], theirs[UNROLL];
        uint4 *remote[UNROLL];
        uint4 a[UNROLL], b[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            int j;
            locate(t + u * stride, mine[u], theirs[u], j);
            remote[u] = reinterpret_cast<uint4 *>(ep.peer[j]);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) b[u] = remote[u][theirs[u]];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) a[u] = local[mine[u]];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            local[mine[u]] = b[u];
            remote[u][theirs[u]] = a[u];
        }
    }
    for (; t < n_items; t += stride) {
        uint64_t mine, theirs;
        int j;
        locate(t, mine, theirs, j);
        uint4 *remote = reinterpret_cast<uint4 *>(ep.peer[j]);
        const uint4 a = local[mine];
        const uint4 b = remote[theirs];
        local[mine] = b;
        remote[theirs] = a;
    }
}

/* Push variant: every rank WRITES its outgoing amplitudes straight into the spare buffers of the
 * ranks that will own them (and copies what it keeps into its own spare buffer); afterwards every
 * rank flips to its spare buffer.  Write-only NVLink traffic (posted stores, no read round trips),
 * one 16-byte local load and one store per unit, and almost no index arithmetic: the destination
 * index is the source index with the victim bits replaced by this rank's selector bits — a
 * constant — and the destination rank is the source index's victim bits.  Needs the spare buffer
 * (twice the shard in memory); exchange_p2p_kernel above is the in-place fallback for shards that
 * fill the GPU (QFT-35 on 4 GPUs: 128 GiB per shard). */
__global__ void __launch_bounds__(256)
exchange_push_kernel(const uint4 *__restrict__ local, const __grid_constant__ ExchangeParams ep) {
    const int k = ep.k;
    const uint64_t n_units = 1ull << ep.n_unit_bits;
    uint64_t vmask = 0, sbits = 0;
    for (int i = 0; i < k; ++i) {
        vmask |= 1ull << ep.victim[i];
        sbits |= (uint64_t)((ep.my_sel >> i) & 1) << ep.victim[i];
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    constexpr int UNROLL = 8;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; t + (UNROLL - 1) * stride < n_units; t += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(local + t + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = t + u * stride;
            int j = 0;
#pragma unroll
            for (int b = 0; b < QGB_MAX_EXCHANGE_LANES; ++b)
                if (b < k) j |= (int)((i >> ep.victim[b]) & 1ull) << b;
            uint4 *dst = reinterpret_cast<uint4 *>(ep.peer[j]);
            __stcs(dst + ((i & ~vmask) | sbits), v[u]);
        }
    }
    for (; t < n_units; t += stride) {
        int j = 0;
        for (int b = 0; b < k; ++b) j |= (int)((t >> ep.victim[b]) & 1ull) << b;
        uint4 *dst = reinterpret_cast<uint4 *>(ep.peer[j]);
        dst[(t & ~vmask) | sbits] = local[t];
    }
}

} // namespace

cudaError_t launch_exchange_push(const void *local, const ExchangeParams &ep, int sm_count, cudaStream_t stream) {
    const uint64_t n_units = 1ull << ep.n_unit_bits;
    uint64_t blocks = (n_units + 256 * 8 - 1) / (256 * 8);
    const uint64_t cap = (uint64_t)(sm_count > 0 ? sm_count : 148) * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    exchange_push_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(local), ep);
    return cudaGetLastError();
}

cudaError_t launch_exchange_p2p(void *local, const ExchangeParams &ep, int sm_count, cudaStream_t stream) {
    const int rest_bits = ep.n_unit_bits - ep.k - 1;
    const uint64_t n_items = (1ull << rest_bits) * ((1ull << ep.k) - 1ull);
    uint64_t blocks = (n_items + 255) / 256;
    const uint64_t cap = (uint64_t)(sm_count > 0 ? sm_count : 148) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    exchange_p2p_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<uint4 *>(local), ep);
    return cudaGetLastError();
}

} // namespace qgb
