// Hypothesis-sharded estimate across the GPUs of one NVLink / NVSwitch box without a collective library on
// the data path (single-pair case).  Each rank scores its slice; the exchange step of the
// path is one 8-byte (count, index) key per pair, so instead of an all-reduce call every rank PUSHES its key
// into every peer's exchange buffer with a system-scope atomicMax over peer memory (NVLink P2P stores /
// atomics), bumps the peer's arrival counter, waits for its own counter to reach `world` and regenerates
// the winning E from the index.  Two tiny kernels after the scoring kernel, no host round trip, no NCCL.
//
// Exchange buffer (one cudaMalloc per rank, shared through CUDA IPC): two slots indexed by call parity,
//   keys[2][pairs] uint64, arrive[2] uint32.
// Slot p = call & 1 is zeroed by its owner right after it consumed it; a peer can write slot p of call c + 2
// only after finishing call c + 1, which needs the owner's arrival for c + 1, which the owner sends after
// that zeroing (stream order), so a slot is always clean when the first key of a call lands in it.
// The wait is bounded (about two seconds of SM clock): on timeout the rank raises a flag, proceeds with what
// has arrived and sfmb200_mg_status reports it - a lost peer must not hang the GPU.
#include "internal.cuh"

namespace sfmb200 {

__global__ void mg_publish_kernel(DeviceState s, MgPeers peers, int parity) {
    // keys first, then (after a system-scope fence) the arrival tickets
    for (int r = 0; r < peers.world; r++) {
        unsigned long long* keys = peers.base[r] + (size_t)parity * s.B;
        for (int b = threadIdx.x; b < s.B; b += blockDim.x) {
            unsigned long long v = s.best[b];
            if (v != 0ull) atomicMax_system(keys + b, v);
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int r = 0; r < peers.world; r++) {
            unsigned int* arrive = reinterpret_cast<unsigned int*>(peers.base[r] + 2 * (size_t)s.B) + parity;
            atomicAdd_system(arrive, 1u);
        }
    }
}

__global__ void mg_collect_kernel(DeviceState s, MgPeers peers, int parity, long long timeout_cycles, int* status) {
    unsigned long long* keys = peers.base[peers.rank] + (size_t)parity * s.B;
    volatile unsigned int* arrive = reinterpret_cast<unsigned int*>(peers.base[peers.rank] + 2 * (size_t)s.B) + parity;
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
    for (int b = threadIdx.x; b < s.B; b += blockDim.x) {
        s.best[b] = *reinterpret_cast<volatile unsigned long long*>(keys + b);
        keys[b] = 0ull;                          // clean for call + 2
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *arrive = 0u;
        if (timed_out) atomicAdd(status, 1);
    }
}

void launch_mg_exchange(const DeviceState& s, const MgPeers& peers, int parity, long long timeout_cycles, int* d_status,
                        cudaStream_t st) {
    mg_publish_kernel<<<1, 128, 0, st>>>(s, peers, parity);
    mg_collect_kernel<<<1, 128, 0, st>>>(s, peers, parity, timeout_cycles, d_status);
}

}  // namespace sfmb200
