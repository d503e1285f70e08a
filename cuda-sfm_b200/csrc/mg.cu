// Hypothesis-sharded estimate across the GPUs of one NVLink / NVSwitch box without a collective library on
// the data path (single-pair case).  Each rank scores its slice; the exchange step of the
// path is one 8-byte (count, index) key per pair, so instead of an all-reduce call every rank PUSHES its key
// into every peer's exchange buffer with a system-scope atomicMax over peer memory (NVLink P2P stores /
// atomics), bumps the peer's arrival counter, waits for its own counter to reach `world` and picks the
// winning E out of the slot of the rank that owns the winning index.  Two tiny kernels after the scoring kernel, no host round trip, no NCCL.
//
// Exchange buffer (one cudaMalloc per rank, shared through CUDA IPC): two slots indexed by call parity,
//   keys[2][pairs] uint64, arrive[2] uint32, E[2][world][pairs][9] float (every rank also pushes the E of its
//   local winner into its own slot, so the global winner's E is already there: nothing is regenerated).
// Slot p = call & 1 is zeroed by its owner right after it consumed it; a peer can write slot p of call c + 2
// only after finishing call c + 1, which needs the owner's arrival for c + 1, which the owner sends after
// that zeroing (stream order), so a slot is always clean when the first key of a call lands in it.
// The wait is bounded (about two seconds of SM clock): on timeout the rank raises a flag, proceeds with what
// has arrived and sfmb200_mg_status reports it - a lost peer must not hang the GPU.
#include "internal.cuh"

namespace sfmb200 {

// Layout of one rank's exchange buffer (in 8-byte words from base): keys[2][B], then arrive[2] (u32) padded to
// 16 bytes, then E[2][world][B][9] floats: slot [parity][r] holds the E of rank r's local winner.
__device__ __forceinline__ unsigned int* mg_arrive(unsigned long long* base, int B, int parity) {
    return reinterpret_cast<unsigned int*>(base + 2 * (size_t)B) + parity;
}
__device__ __forceinline__ float* mg_E(unsigned long long* base, int B, int world, int parity, int r) {
    return reinterpret_cast<float*>(base + 2 * (size_t)B + 2) + (((size_t)parity * world + r) * B) * 9;
}

__global__ void mg_publish_kernel(DeviceState s, MgPeers peers, int parity) {
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
        for (int r = 0; r < peers.world; r++) atomicAdd_system(mg_arrive(peers.base[r], s.B, parity), 1u);
}

__global__ void mg_collect_kernel(DeviceState s, MgPeers peers, int parity, int H_total, long long timeout_cycles, int* status) {
    unsigned long long* base = peers.base[peers.rank];
    unsigned long long* keys = base + (size_t)parity * s.B;
    volatile unsigned int* arrive = mg_arrive(base, s.B, parity);
    __shared__ int timed_out;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        timed_out = 0;
        while (*arrive < (unsigned int)peers.world) {
            if (clock64() - t0 > timeout_cycles) { timed_out = 1; break; }
            __nanosleep(200);
        }
        __threadfence_system();
    }
    __syncthreads();
    const long long per = H_total / peers.world, rem = H_total % peers.world;
    for (int b = threadIdx.x; b < s.B; b += blockDim.x) {
        const unsigned long long packed = *reinterpret_cast<volatile unsigned long long*>(keys + b);
        keys[b] = 0ull;                          // clean for call + 2
        s.best[b] = packed;
        const unsigned int hg = 0xFFFFFFFFu - (unsigned int)(packed & 0xFFFFFFFFull);
        s.best_idx[b] = (int)hg;
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
    __syncthreads();
    if (threadIdx.x == 0) {
        *arrive = 0u;
        if (timed_out) atomicAdd(status, 1);
    }
}

size_t mg_buffer_bytes(int B, int world) {
    return (2 * (size_t)B + 2) * sizeof(unsigned long long) + 2 * (size_t)world * B * 9 * sizeof(float);
}

void launch_mg_exchange(const DeviceState& s, const MgPeers& peers, int parity, int H_total, long long timeout_cycles, int* d_status,
                        cudaStream_t st) {
    mg_publish_kernel<<<1, 128, 0, st>>>(s, peers, parity);
    mg_collect_kernel<<<1, 128, 0, st>>>(s, peers, parity, H_total, timeout_cycles, d_status);
}

}  // namespace sfmb200
