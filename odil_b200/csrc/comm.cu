// Slab halo exchange and scalar all-reduce over NVLink peer memory.
//
// One process per GPU.  Every rank owns ONE cudaMalloc'ed block -- flags, reduction slots and halo staging -- that it
// exports with cudaIpcGetMemHandle; the host plumbing (torch.distributed) all-gathers the 64-byte handles and every
// rank maps the blocks of its peers (cudaIpcOpenMemHandle -> NVLink / NVSwitch peer access).  After that the data
// plane is two kernels per exchange and no host involvement:
//
//   k_halo_push   copies this rank's boundary planes of all arrays straight into the neighbours' staging areas
//                 (16-byte peer stores over NVLink) and, when the last block is done, publishes the sequence number in
//                 the neighbours' flag words (release, system scope);
//   k_halo_pull   waits for the two neighbours' flags (acquire), copies staging -> halo planes, advances the sequence.
//
// Staging and flags are double-buffered by sequence parity: a rank can be at most one exchange ahead of its
// neighbour (its own pull of exchange s-1 needs the neighbour's push of s-1, which the neighbour issues after its pull
// of s-2), so the area written for exchange s was consumed two exchanges ago.  The sequence counter lives in device
// memory, hence a captured CUDA graph of the whole epoch replays without patching.  Compared with an NCCL send/recv
// group (~40 us per group on this box, five groups per epoch in round 1) an exchange costs two small launches.
//
// odil_b200_allreduce_scalars: every rank stores its K <= 16 doubles into every peer's slot, waits for all flags and
// sums in rank order -- one kernel, deterministic, identical on all ranks.
#include <cstring>
#include <vector>

#include "common.cuh"

namespace odil {

constexpr int kMaxWorld = 16;
constexpr int kMaxRed = 16;
constexpr int kMaxItems = 40;

struct alignas(128) FlagLine {
    unsigned int v;
    unsigned int pad[31];
};

// Start of every rank's exported block.
struct CommShared {
    FlagLine halo_flag[2][2];          // [direction: 0 = written by the lower neighbour, 1 = by the upper][parity]
    FlagLine red_flag[kMaxWorld][2];   // [source rank][parity]
    double red_slot[kMaxWorld][2][kMaxRed];
};

// Rank-private device state.
struct CommLocal {
    unsigned int halo_seq;
    unsigned int red_seq;
    unsigned int push_done;
    unsigned int pull_done;
};

struct HaloItem {
    const char* send_lo;   // first owned planes  -> lower neighbour's upper halo
    const char* send_hi;   // last owned planes   -> upper neighbour's lower halo
    char* recv_lo;         // lower halo planes
    char* recv_hi;         // upper halo planes
    long long nbytes;      // multiple of `vec`
    long long off;         // offset inside a staging area (256-byte aligned)
    int vec;               // bytes per access: 16 when pointers and size allow, else 8 or 4
};

struct HaloParams {
    int nitems;
    long long stage_bytes;       // capacity of one staging area
    char* my_base;
    char* lo_base;               // lower neighbour's block
    char* hi_base;               // upper neighbour's block
    CommLocal* loc;
    HaloItem it[kMaxItems];
};

__device__ __forceinline__ char* stage_ptr(char* base, long long stage_bytes, int dir, int parity) {
    return base + sizeof(CommShared) + (long long)(dir * 2 + parity) * stage_bytes;
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <typename V>
__device__ __forceinline__ void copy_as(char* __restrict__ dst, const char* __restrict__ src, long long nbytes, long long tid,
                                        long long nthreads) {
    const V* s = reinterpret_cast<const V*>(src);
    V* d = reinterpret_cast<V*>(dst);
    const long long n = nbytes / (long long)sizeof(V);
    for (long long i = tid; i < n; i += nthreads) d[i] = s[i];
}

__device__ __forceinline__ void copy16(char* __restrict__ dst, const char* __restrict__ src, long long nbytes, int vec,
                                       long long tid, long long nthreads) {
    if (vec == 16)
        copy_as<int4>(dst, src, nbytes, tid, nthreads);
    else if (vec == 8)
        copy_as<long long>(dst, src, nbytes, tid, nthreads);
    else
        copy_as<int>(dst, src, nbytes, tid, nthreads);
}

static __global__ void __launch_bounds__(256) k_halo_push(const HaloParams p) {
    const unsigned int s = p.loc->halo_seq + 1;
    const int par = s & 1;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    char* to_lo = stage_ptr(p.lo_base, p.stage_bytes, 1, par);  // I am the lower neighbour's UPPER neighbour
    char* to_hi = stage_ptr(p.hi_base, p.stage_bytes, 0, par);
    for (int i = 0; i < p.nitems; ++i) {
        copy16(to_lo + p.it[i].off, p.it[i].send_lo, p.it[i].nbytes, p.it[i].vec, tid, nth);
        copy16(to_hi + p.it[i].off, p.it[i].send_hi, p.it[i].nbytes, p.it[i].vec, tid, nth);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&p.loc->push_done, 1u);
        if (t == gridDim.x - 1) {
            p.loc->push_done = 0;
            __threadfence_system();
            st_release_sys(&reinterpret_cast<CommShared*>(p.lo_base)->halo_flag[1][par].v, s);
            st_release_sys(&reinterpret_cast<CommShared*>(p.hi_base)->halo_flag[0][par].v, s);
        }
    }
}

static __global__ void __launch_bounds__(256) k_halo_pull(const HaloParams p) {
    const unsigned int s = p.loc->halo_seq + 1;
    const int par = s & 1;
    CommShared* me = reinterpret_cast<CommShared*>(p.my_base);
    if (threadIdx.x == 0) {
        while ((int)(ld_acquire_sys(&me->halo_flag[0][par].v) - s) < 0) {
        }
        while ((int)(ld_acquire_sys(&me->halo_flag[1][par].v) - s) < 0) {
        }
    }
    __syncthreads();
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const char* from_lo = stage_ptr(p.my_base, p.stage_bytes, 0, par);
    const char* from_hi = stage_ptr(p.my_base, p.stage_bytes, 1, par);
    for (int i = 0; i < p.nitems; ++i) {
        copy16(p.it[i].recv_lo, from_lo + p.it[i].off, p.it[i].nbytes, p.it[i].vec, tid, nth);
        copy16(p.it[i].recv_hi, from_hi + p.it[i].off, p.it[i].nbytes, p.it[i].vec, tid, nth);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&p.loc->pull_done, 1u);
        if (t == gridDim.x - 1) {
            p.loc->pull_done = 0;
            p.loc->halo_seq = s;
        }
    }
}

// Consumer of a halo ACCUMULATION: what the lower / upper neighbour pushed is ADDED to this rank's first / last owned
// planes (recv_lo / recv_hi point at them).  Used for the push-style transposed interpolation: every rank computes
// partial sums for the coarse planes one beyond its slab and hands them to their owners once per epoch.
template <typename T>
static __global__ void __launch_bounds__(256) k_halo_pull_add(const HaloParams p) {
    const unsigned int s = p.loc->halo_seq + 1;
    const int par = s & 1;
    CommShared* me = reinterpret_cast<CommShared*>(p.my_base);
    if (threadIdx.x == 0) {
        while ((int)(ld_acquire_sys(&me->halo_flag[0][par].v) - s) < 0) {
        }
        while ((int)(ld_acquire_sys(&me->halo_flag[1][par].v) - s) < 0) {
        }
    }
    __syncthreads();
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const char* from_lo = stage_ptr(p.my_base, p.stage_bytes, 0, par);
    const char* from_hi = stage_ptr(p.my_base, p.stage_bytes, 1, par);
    for (int i = 0; i < p.nitems; ++i) {
        const long long n = p.it[i].nbytes / (long long)sizeof(T);
        const T* a = reinterpret_cast<const T*>(from_lo + p.it[i].off);
        const T* b = reinterpret_cast<const T*>(from_hi + p.it[i].off);
        T* lo = reinterpret_cast<T*>(p.it[i].recv_lo);
        T* hi = reinterpret_cast<T*>(p.it[i].recv_hi);
        // fixed order: first the lower neighbour's contribution, then the upper one's (they may hit the same plane
        // when a slab is one plane thick) -> deterministic sums
        if (lo == hi) {
            for (long long e = tid; e < n; e += nth) lo[e] = (lo[e] + a[e]) + b[e];
        } else {
            for (long long e = tid; e < n; e += nth) {
                lo[e] += a[e];
                hi[e] += b[e];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&p.loc->pull_done, 1u);
        if (t == gridDim.x - 1) {
            p.loc->pull_done = 0;
            p.loc->halo_seq = s;
        }
    }
}

struct RedParams {
    int rank, world, count;
    char* base[kMaxWorld];
    CommLocal* loc;
    double* vals;
};

static __global__ void __launch_bounds__(32) k_allreduce(const RedParams p) {
    const unsigned int s = p.loc->red_seq + 1;
    const int par = s & 1;
    const int t = threadIdx.x;
    if (t < p.count) {
        const double v = p.vals[t];
        for (int r = 0; r < p.world; ++r) reinterpret_cast<CommShared*>(p.base[r])->red_slot[p.rank][par][t] = v;
    }
    __threadfence_system();
    __syncwarp();
    if (t < p.world) st_release_sys(&reinterpret_cast<CommShared*>(p.base[t])->red_flag[p.rank][par].v, s);
    CommShared* me = reinterpret_cast<CommShared*>(p.base[p.rank]);
    if (t < p.world) {
        while ((int)(ld_acquire_sys(&me->red_flag[t][par].v) - s) < 0) {
        }
    }
    __syncwarp();
    if (t < p.count) {
        double acc = 0.0;
        for (int r = 0; r < p.world; ++r) acc += me->red_slot[r][par][t];  // rank order: same bits on every rank
        p.vals[t] = acc;
    }
    __syncwarp();
    if (t == 0) p.loc->red_seq = s;
}

}  // namespace odil

using namespace odil;

struct odil_b200_comm {
    int rank = 0, world = 1;
    long long stage_bytes = 0;
    char* block = nullptr;           // this rank's exported block
    size_t block_bytes = 0;
    CommLocal* loc = nullptr;
    char* peer[kMaxWorld] = {nullptr};
    bool opened[kMaxWorld] = {false};
    bool connected = false;
};

extern "C" {

int odil_b200_comm_create(int rank, int world, int64_t halo_bytes, odil_b200_comm** comm, void* handle_out) {
    ODIL_REQUIRE(comm && handle_out, "null argument");
    ODIL_REQUIRE(world >= 2 && world <= kMaxWorld && rank >= 0 && rank < world, "rank=%d world=%d out of range", rank, world);
    ODIL_REQUIRE(halo_bytes >= 0, "negative staging size");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    odil_b200_comm* c = new odil_b200_comm;
    c->rank = rank;
    c->world = world;
    c->stage_bytes = ((halo_bytes + 255) / 256) * 256;
    c->block_bytes = sizeof(CommShared) + (size_t)4 * c->stage_bytes;
    if (cudaMalloc(&c->block, c->block_bytes) != cudaSuccess || cudaMalloc(&c->loc, sizeof(CommLocal)) != cudaSuccess) {
        delete c;
        return fail("cudaMalloc of the %zu-byte communication block failed: %s", c->block_bytes,
                    cudaGetErrorString(cudaGetLastError()));
    }
    ODIL_CUDA(cudaMemset(c->block, 0, sizeof(CommShared)));
    ODIL_CUDA(cudaMemset(c->loc, 0, sizeof(CommLocal)));
    ODIL_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    ODIL_CUDA(cudaIpcGetMemHandle(&h, c->block));
    std::memcpy(handle_out, &h, sizeof(h));
    c->peer[rank] = c->block;
    *comm = c;
    return 0;
}

// handles: world x 64 bytes in rank order (this rank's own entry is ignored).
int odil_b200_comm_connect(odil_b200_comm* c, const void* handles) {
    ODIL_REQUIRE(c && handles, "null argument");
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return fail("cudaIpcOpenMemHandle(rank %d) failed: %s (slab runs need NVLink/PCIe peer access between the GPUs "
                        "of one node)", r, cudaGetErrorString(e));
        c->peer[r] = (char*)p;
        c->opened[r] = true;
    }
    c->connected = true;
    return 0;
}

int64_t odil_b200_comm_capacity(const odil_b200_comm* c) { return c ? c->stage_bytes : 0; }

// Ring exchange of `narrays` arrays: send_lo[i] / send_hi[i] point at this rank's first / last `nbytes[i]` owned bytes,
// recv_lo[i] / recv_hi[i] at its lower / upper halo.  Replaces the batched ncclSend/ncclRecv group of round 1
// (odil_b200/slab.py exchange).  Asynchronous on `stream`; capturable.
static int halo_run(odil_b200_comm* c, int narrays, const void* const* send_lo, const void* const* send_hi,
                    void* const* recv_lo, void* const* recv_hi, const int64_t* nbytes, int add_dtype, void* stream) {
    ODIL_REQUIRE(c && c->connected, "communicator is not connected");
    ODIL_REQUIRE(narrays >= 1 && narrays <= kMaxItems, "narrays=%d out of range (1..%d)", narrays, kMaxItems);
    HaloParams p;
    p.nitems = narrays;
    p.stage_bytes = c->stage_bytes;
    p.my_base = c->block;
    p.lo_base = c->peer[(c->rank + c->world - 1) % c->world];
    p.hi_base = c->peer[(c->rank + 1) % c->world];
    p.loc = c->loc;
    long long off = 0, total16 = 0;
    for (int i = 0; i < narrays; ++i) {
        const uintptr_t bits = (uintptr_t)send_lo[i] | (uintptr_t)send_hi[i] | (uintptr_t)recv_lo[i] | (uintptr_t)recv_hi[i] |
                               (uintptr_t)nbytes[i];
        ODIL_REQUIRE(nbytes[i] > 0 && bits % 4 == 0, "array %d: halo planes must be 4-byte aligned and sized (%lld bytes)", i,
                     (long long)nbytes[i]);
        const int vec = bits % 16 == 0 ? 16 : (bits % 8 == 0 ? 8 : 4);
        p.it[i] = HaloItem{(const char*)send_lo[i], (const char*)send_hi[i], (char*)recv_lo[i], (char*)recv_hi[i],
                           (long long)nbytes[i], off, vec};
        off += (nbytes[i] + 255) / 256 * 256;
        total16 += nbytes[i] >> 4;
    }
    ODIL_REQUIRE(off <= c->stage_bytes, "exchange of %lld bytes exceeds the staging capacity %lld", off, c->stage_bytes);
    // Peer stores have microseconds of latency: bandwidth comes from the number of 16-byte stores in flight, so the
    // copy is spread over up to four CTAs per SM (64 CTAs moved 5.4 MB in ~40 us at 512^3; NVLink can do it in ~8).
    int blocks = (int)std::min<long long>(148 * 4, std::max<long long>(1, (2 * total16 + 511) / 512));
    cudaStream_t st = (cudaStream_t)stream;
    k_halo_push<<<blocks, 256, 0, st>>>(p);
    ODIL_LAUNCHED();
    if (add_dtype == ODIL_B200_F32)
        k_halo_pull_add<float><<<blocks, 256, 0, st>>>(p);
    else if (add_dtype == ODIL_B200_F64)
        k_halo_pull_add<double><<<blocks, 256, 0, st>>>(p);
    else
        k_halo_pull<<<blocks, 256, 0, st>>>(p);
    ODIL_LAUNCHED();
    return 0;
}

int odil_b200_halo_exchange(odil_b200_comm* c, int narrays, const void* const* send_lo, const void* const* send_hi,
                            void* const* recv_lo, void* const* recv_hi, const int64_t* nbytes, void* stream) {
    return halo_run(c, narrays, send_lo, send_hi, recv_lo, recv_hi, nbytes, -1, stream);
}

// Ring ACCUMULATION: send_lo[i] / send_hi[i] = this rank's partial sums for the plane(s) just below / above its slab
// (owned by the lower / upper neighbour); acc_lo[i] / acc_hi[i] = this rank's first / last owned plane(s), to which the
// lower / upper neighbour's partial sums are added.  dtype: element type of every array.
int odil_b200_halo_accumulate(odil_b200_comm* c, int narrays, const void* const* send_lo, const void* const* send_hi,
                              void* const* acc_lo, void* const* acc_hi, const int64_t* nbytes, int dtype, void* stream) {
    ODIL_REQUIRE(dtype == ODIL_B200_F32 || dtype == ODIL_B200_F64, "dtype=%d unsupported", dtype);
    return halo_run(c, narrays, send_lo, send_hi, acc_lo, acc_hi, nbytes, dtype, stream);
}

// In-place sum over all ranks of `count` doubles in device memory (loss terms, dot products).  Replaces ncclAllReduce.
int odil_b200_allreduce_scalars(odil_b200_comm* c, double* dev_scalars, int count, void* stream) {
    ODIL_REQUIRE(c && c->connected, "communicator is not connected");
    ODIL_REQUIRE(count >= 1 && count <= kMaxRed, "count=%d out of range (1..%d)", count, kMaxRed);
    RedParams p;
    p.rank = c->rank;
    p.world = c->world;
    p.count = count;
    for (int r = 0; r < c->world; ++r) p.base[r] = c->peer[r];
    p.loc = c->loc;
    p.vals = dev_scalars;
    k_allreduce<<<1, 32, 0, (cudaStream_t)stream>>>(p);
    ODIL_LAUNCHED();
    return 0;
}

int odil_b200_comm_destroy(odil_b200_comm* c) {
    if (!c) return 0;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->block) cudaFree(c->block);
    if (c->loc) cudaFree(c->loc);
    delete c;
    return 0;
}

}  // extern "C"
