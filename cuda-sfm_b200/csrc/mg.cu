// Hypothesis-sharded estimate across the GPUs of one NVLink / NVSwitch box without a collective library on
// the data path (single-pair case).  Each rank scores its slice; the exchange step of the
// path is one 8-byte (count, index) key per pair, so instead of an all-reduce call every rank PUSHES its key
// into every peer's exchange buffer with a system-scope atomicMax over peer memory (NVLink P2P stores /
// atomics), bumps the peer's arrival counter, waits for its own counter to show that every rank has published
// this call and picks the winning E out of the slot of the rank that owns the winning index.  Two tiny kernels
// after the scoring kernel, no host round trip, no NCCL.
//
// Exchange buffer (one cudaMalloc per rank, shared through CUDA IPC), two slots indexed by call parity:
//   keys[2][pairs] uint64, arrive uint64 (ONE monotone counter), E[2][world][pairs][9] float (every rank also
//   pushes the E of its local winner into its own slot, so the global winner's E is already there: nothing is
//   regenerated).
// Call number c (0, 1, 2, ...) uses slot c & 1 and is complete on a rank when arrive >= world * (c + 1): the
// counter is never reset, so an arrival can never be mistaken for one of another call.  The owner zeroes the
// key slot right after consuming it; a peer can write slot c & 1 again (call c + 2) only after finishing call
// c + 1, which needs the owner's arrival for c + 1, which the owner sends after that zeroing (stream order), so
// a slot is always clean when the first key of a call lands in it.
// The wait is bounded (about two seconds of SM clock by default).  A wait that times out POISONS the exchange:
// the rank raises its sticky `failed` word (pinned host memory, visible to the host without a synchronise),
// publishes no result (count 0, E = 0) for that and every later call, and sfmb200_estimate_e_mg returns an
// error from the next call on until the job reconnects (sfmb200_mg_close + mg_init + mg_connect on every
// rank) - a late peer's key landing in a recycled slot can therefore never be taken for a result.
#include "internal.cuh"

namespace sfmb200 {

// Layout of one rank's exchange buffer (in 8-byte words from base): keys[2][B], then arrive (u64) + one spare
// word, then E[2][world][B][9] floats: slot [parity][r] holds the E of rank r's local winner.
__device__ __forceinline__ unsigned long long* mg_arrive(unsigned long long* base, int B) { return base + 2 * (size_t)B; }
__device__ __forceinline__ float* mg_E(unsigned long long* base, int B, int world, int parity, int r) {
    return reinterpret_cast<float*>(base + 2 * (size_t)B + 2) + (((size_t)parity * world + r) * B) * 9;
}

__global__ void mg_publish_kernel(DeviceState s, MgPeers peers, int parity, const volatile int* failed) {
    if (*failed != 0) return;                    // poisoned: stay silent, the peers time out and poison themselves
    // this rank's local winner: key by system-scope atomicMax, E by plain stores into its own slot of every peer;
    // then (after a system-scope fence) the arrival tickets
    for (int r = 0; r < peers.world; r++) {
        unsigned long long* keys = peers.base[r] + (size_t)parity * s.B;
        float* Es = mg_E(peers.base[r], s.B, peers.world, parity, peers.rank);
        for (int b = threadIdx.x; b < s.B; b += blockDim.x) {
            unsigned long long v = s.best[b];
            if (v != 0ull) atomicMax_system(keys + b, v);
#pragma unroll
            for (int k = 0; k < 9; k++) Es[(size_t)b * 9 + k] = s.E[(size_t)b * 9 + k];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        for (int r = 0; r < peers.world; r++) atomicAdd_system(mg_arrive(peers.base[r], s.B), 1ull);
}

__global__ void mg_collect_kernel(DeviceState s, MgPeers peers, int parity, unsigned long long target, int H_total,
                                  long long timeout_cycles, volatile int* failed) {
    unsigned long long* base = peers.base[peers.rank];
    unsigned long long* keys = base + (size_t)parity * s.B;
    volatile unsigned long long* arrive = mg_arrive(base, s.B);
    __shared__ int bad;
    if (threadIdx.x == 0) {
        bad = *failed != 0;
        const long long t0 = clock64();
        while (!bad && *arrive < target) {
            if (clock64() - t0 > timeout_cycles) { bad = 1; break; }
            __nanosleep(200);
        }
        __threadfence_system();
    }
    __syncthreads();
    const long long per = H_total / peers.world, rem = H_total % peers.world;
    for (int b = threadIdx.x; b < s.B; b += blockDim.x) {
        const unsigned long long packed = bad ? 0ull : *reinterpret_cast<volatile unsigned long long*>(keys + b);
        keys[b] = 0ull;                          // clean for call + 2
        s.best[b] = packed;
        const unsigned int hg = 0xFFFFFFFFu - (unsigned int)(packed & 0xFFFFFFFFull);
        s.best_idx[b] = packed != 0ull ? (int)hg : -1;
        s.best_count[b] = (int)(packed >> 32);
        // the slice that owns hypothesis hg (same split as sfmb200_estimate_e_mg / sharding.shard_range)
        int owner = 0;
        if (packed != 0ull) {
            const long long h = (long long)hg, big = (per + 1) * rem;
            owner = (int)(h < big ? h / (per + 1) : rem + (h - big) / (per > 0 ? per : 1));
            if (owner >= peers.world) owner = peers.world - 1;
        }
        const volatile float* Es = mg_E(base, s.B, peers.world, parity, owner) + (size_t)b * 9;
#pragma unroll
        for (int k = 0; k < 9; k++) s.E[(size_t)b * 9 + k] = packed != 0ull ? Es[k] : 0.0f;
    }
    if (threadIdx.x == 0 && bad) {
        *failed = *failed + 1;                   // sticky; counts the calls that produced no result
        __threadfence_system();
    }
}

size_t mg_buffer_bytes(int B, int world) {
    return (2 * (size_t)B + 2) * sizeof(unsigned long long) + 2 * (size_t)world * B * 9 * sizeof(float);
}

void launch_mg_exchange(const DeviceState& s, const MgPeers& peers, long long call, int H_total, long long timeout_cycles,
                        int* d_failed, cudaStream_t st) {
    const int parity = (int)(call & 1);
    mg_publish_kernel<<<1, 128, 0, st>>>(s, peers, parity, d_failed);
    mg_collect_kernel<<<1, 128, 0, st>>>(s, peers, parity, (unsigned long long)peers.world * (unsigned long long)(call + 1), H_total,
                                         timeout_cycles, d_failed);
}

}  // namespace sfmb200
