// Gradient all-reduce over NVLink / NVSwitch peer memory (the one exchange step of the data-parallel
// path; precedent legacy/train_tpu.py:115 `xm.optimizer_step`). One process per GPU; every rank's flat gradient buffer
// and flag pad live in symmetric memory, so each rank holds device pointers to all peers' copies.
//
// Why not NCCL for this: its all-reduce kernels need SM slots of their own (large blocks + shared memory) and cannot
// co-reside with the persistent 1-CTA-per-SM convolution kernels of the backward pass (~200 KB of shared memory each), so a
// bucket launched "in the background" waits for a conv kernel to end and then keeps some SMs away from the next one,
// whose static tile schedule stretches by that delay (measured: +0.5 ms / step at 2 GPUs). This kernel
// uses NO shared memory, 256 threads and < 48 registers: its CTAs fit beside a convolution CTA on the same SM -- like the
// BatchNorm passes that already overlap the weight gradients -- and it moves the bytes with plain peer loads / stores.
//
// Algorithm per bucket (in place, deterministic, identical bits on every rank):
//   rank r owns slice r of the bucket. It (1) tells every peer "my copy of this bucket is final" (flag write, release at
//   system scope), (2) waits for the same message from every peer, (3) for each element of ITS slice loads the value from
//   all ranks in rank order, sums, scales by 1/world and stores the result into the slice on ALL ranks, (4) after a
//   system-scope fence its last CTA tells every peer "slice r of this bucket has arrived". cvb_allreduce_wait makes the
//   consumer stream wait for those messages from all ranks. Slice r is read and written by rank r only, so nothing
//   races; flags carry the step's epoch (monotone), so they are never reset.
#include "common.cuh"

namespace cvb {

constexpr int kArThreads = 256;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Peer DATA moves with ordinary (weak) 16-byte loads and stores, like any elementwise kernel: measured 660-760 GB/s
// per direction through these mappings on a 2 x B200 NV18 box. Ordering comes from the flags alone: st.release.sys /
// ld.acquire.sys around them (the acquire also invalidates this SM's L1: CCTL.IVALL in the SASS) and a system-scope fence
// between the last data store and the "arrived" flag. Strong accesses for the data are a trap: ld/st.relaxed.sys AND
// __ldcg/__stcg (LDG/STG.STRONG.GPU) both ran at ~43 GB/s on peer memory -- 3.4 ms for the 138 MB of UNet gradients.
__device__ __forceinline__ float4 ld_peer_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_peer_f4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float ld_peer_f1(const float* p) { return *p; }
__device__ __forceinline__ void st_peer_f1(float* p, float v) { *p = v; }

struct CommDev {
  float* bufs[CVB_COMM_MAX_WORLD];
  uint32_t* flags[CVB_COMM_MAX_WORLD];
  float* mc;  // multicast (NVLS) address of the gradient buffer, or NULL
  int rank, world;
};

// NVLink SHARP: one load returns the SUM over every rank's copy (reduced inside the NVSwitch), one store writes every
// rank's copy. fp32 accumulation; every element is reduced once, by its slice's owner, and broadcast: identical bits on
// all ranks.
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// flag pad layout (uint32): [0, MAXB*MAXW) = "ready" [bucket][rank]; [MAXB*MAXW, 2*MAXB*MAXW) = "arrived" [bucket][rank];
// the last word = CTA ticket of the local kernel
__device__ __forceinline__ int ready_idx(int b, int r) { return b * CVB_COMM_MAX_WORLD + r; }
__device__ __forceinline__ int done_idx(int b, int r) { return (CVB_COMM_MAX_BUCKETS + b) * CVB_COMM_MAX_WORLD + r; }
constexpr int kTicketIdx = 2 * CVB_COMM_MAX_BUCKETS * CVB_COMM_MAX_WORLD;

__device__ __forceinline__ long long slice_bound(long long offset, long long count, int k, int world) {
  if (k <= 0) return offset;
  if (k >= world) return offset + count;
  const long long b = offset + count * k / world;
  const long long a = (b + 3) & ~3LL;  // slices start on 16-byte boundaries of the (16-byte aligned) buffer
  return a < offset + count ? a : offset + count;
}

template <int WORLD>  // compile-time rank count: WORLD x U independent 16-byte loads in flight per thread
__global__ void __launch_bounds__(kArThreads) allreduce_mean_kernel(CommDev c, long long offset, long long count,
                                                                     int bucket, uint32_t epoch, int dbg) {
  uint32_t* my_flags = c.flags[c.rank];
  // (1) my bucket is final: every kernel that wrote it precedes this one in stream order
  if (blockIdx.x == 0 && threadIdx.x < c.world) {
    __threadfence_system();
    st_release_sys(c.flags[threadIdx.x] + ready_idx(bucket, c.rank), epoch);
  }
  // (2) every peer's bucket is final
  if (threadIdx.x < c.world) {
    const uint32_t* f = my_flags + ready_idx(bucket, threadIdx.x);
    while (ld_acquire_sys(f) != epoch) __nanosleep(64);
  }
  __syncthreads();
  // (3) my slice
  const long long lo = slice_bound(offset, count, c.rank, c.world), hi = slice_bound(offset, count, c.rank + 1, c.world);
  const float inv = 1.f / static_cast<float>(c.world);
  const long long v_lo = (lo + 3) & ~3LL, v_hi = hi & ~3LL;  // vector body; a ragged head / tail goes element by element
  if (v_hi > v_lo) {
    const long long n4 = (v_hi - v_lo) >> 2;
    const long long stride = 1LL * gridDim.x * kArThreads;
    if (c.mc != nullptr) {
      // NVLS path: 1 load + 1 store per 16 bytes whatever the rank count (instead of WORLD + WORLD)
      constexpr int UM = 4;
      for (long long i = 1LL * blockIdx.x * kArThreads + threadIdx.x; i < n4; i += UM * stride) {
        float4 v[UM];
#pragma unroll
        for (int u = 0; u < UM; ++u)
          if (i + u * stride < n4) v[u] = multimem_ld_reduce_f4(c.mc + v_lo + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < UM; ++u)
          if (i + u * stride < n4) {
            v[u].x *= inv; v[u].y *= inv; v[u].z *= inv; v[u].w *= inv;
            multimem_st_f4(c.mc + v_lo + 4 * (i + u * stride), v[u]);
          }
      }
    } else {
    constexpr int U = WORLD <= 2 ? 4 : (WORLD <= 4 ? 2 : 1);  // vectors per thread and iteration
    for (long long i = 1LL * blockIdx.x * kArThreads + threadIdx.x; i < n4; i += U * stride) {
      float4 v[WORLD][U];
      bool on[U];
#pragma unroll
      for (int u = 0; u < U; ++u) on[u] = i + u * stride < n4;
#pragma unroll
      for (int p = 0; p < WORLD; ++p)  // every load first ...
#pragma unroll
        for (int u = 0; u < U; ++u)
          v[p][u] = on[u] ? ld_peer_f4(c.bufs[((dbg & 1) && p != c.rank) ? c.rank : p] + v_lo + 4 * (i + u * stride))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < U; ++u) {  // ... then the sums in rank order: the same bits on every rank, every step
        float4 a = v[0][u];
#pragma unroll
        for (int p = 1; p < WORLD; ++p) {
          a.x += v[p][u].x; a.y += v[p][u].y; a.z += v[p][u].z; a.w += v[p][u].w;
        }
        a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
        v[0][u] = a;
      }
#pragma unroll
      for (int p = 0; p < WORLD; ++p)
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (on[u] && !((dbg & 2) && p != c.rank) && !((dbg & 4) && p == c.rank))
            st_peer_f4(c.bufs[p] + v_lo + 4 * (i + u * stride), v[0][u]);
    }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < 8) {  // ragged head [lo, v_lo) and tail [v_hi, hi): at most 3 elements each
    const long long head = (v_hi > v_lo ? v_lo : hi) - lo;
    const long long e = threadIdx.x < 4 ? lo + threadIdx.x : (v_hi > v_lo ? v_hi : hi) + (threadIdx.x - 4);
    const bool mine = threadIdx.x < 4 ? threadIdx.x < head : e < hi;
    if (mine) {
      float a = 0.f;
      for (int p = 0; p < c.world; ++p) a += ld_peer_f1(c.bufs[p] + e);
      a *= inv;
      for (int p = 0; p < c.world; ++p) st_peer_f1(c.bufs[p] + e, a);
    }
  }
  // (4) slice r has arrived everywhere
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(my_flags + kTicketIdx, 1u);
    last = t == gridDim.x - 1;
    if (last) my_flags[kTicketIdx] = 0u;  // launches of one rank are stream-ordered: the next one starts from 0
  }
  __syncthreads();
  if (last && threadIdx.x < c.world) {
    __threadfence_system();
    st_release_sys(c.flags[threadIdx.x] + done_idx(bucket, c.rank), epoch);
  }
}

__global__ void allreduce_wait_kernel(CommDev c, int n_buckets, uint32_t epoch) {
  const uint32_t* my_flags = c.flags[c.rank];
  for (int i = threadIdx.x; i < n_buckets * c.world; i += blockDim.x) {
    const uint32_t* f = my_flags + done_idx(i / c.world, i % c.world);
    while (ld_acquire_sys(f) != epoch) __nanosleep(64);
  }
}

static int load_comm(const cvb_comm* comm, CommDev* c) {
  CVB_REQUIRE(comm && comm->peer_bufs_host && comm->peer_flags_host, CVB_ERR_INVALID_ARG, "allreduce: null communicator");
  CVB_REQUIRE(comm->world >= 1 && comm->world <= CVB_COMM_MAX_WORLD && comm->rank >= 0 && comm->rank < comm->world,
              CVB_ERR_INVALID_ARG, "allreduce: rank %d / world %d (max %d ranks)", comm->rank, comm->world, CVB_COMM_MAX_WORLD);
  for (int p = 0; p < CVB_COMM_MAX_WORLD; ++p) {
    c->bufs[p] = p < comm->world ? static_cast<float*>(comm->peer_bufs_host[p]) : nullptr;
    c->flags[p] = p < comm->world ? static_cast<uint32_t*>(comm->peer_flags_host[p]) : nullptr;
    if (p < comm->world) {
      CVB_REQUIRE(c->bufs[p] && c->flags[p], CVB_ERR_INVALID_ARG, "allreduce: null peer pointer for rank %d", p);
      CVB_REQUIRE((reinterpret_cast<uintptr_t>(c->bufs[p]) & 15) == 0, CVB_ERR_INVALID_ARG, "allreduce: peer buffer not 16-byte aligned");
    }
  }
  c->mc = static_cast<float*>(comm->multicast_buf);
  CVB_REQUIRE((reinterpret_cast<uintptr_t>(c->mc) & 15) == 0, CVB_ERR_INVALID_ARG, "allreduce: multicast address not 16-byte aligned");
  c->rank = comm->rank;
  c->world = comm->world;
  return CVB_OK;
}

}  // namespace cvb

using namespace cvb;

extern "C" int cvb_comm_flag_words(void) { return kTicketIdx + 4; }

extern "C" int cvb_allreduce_mean_f32(const cvb_comm* comm, int64_t offset, int64_t count, int bucket, uint32_t epoch,
                                      int ctas, void* stream) {
  CommDev c;
  int rc = load_comm(comm, &c);
  if (rc) return rc;
  CVB_REQUIRE(offset >= 0 && count > 0, CVB_ERR_INVALID_ARG, "allreduce: empty range");
  CVB_REQUIRE(bucket >= 0 && bucket < CVB_COMM_MAX_BUCKETS && epoch != 0, CVB_ERR_INVALID_ARG,
              "allreduce: bucket %d (max %d) / epoch %u", bucket, CVB_COMM_MAX_BUCKETS, epoch);
  if (ctas < 1) ctas = 1;
  if (ctas > 4 * sm_count()) ctas = 4 * sm_count();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static int dbg = -1;  // experiment knob CVB_AR_DEBUG: 1 = no remote loads, 2 = no remote stores, 4 = no local stores
  if (dbg < 0) {
    const char* e = getenv("CVB_AR_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  switch (c.world) {
#define CVB_AR_CASE(W) case W: allreduce_mean_kernel<W><<<ctas, kArThreads, 0, st>>>(c, offset, count, bucket, epoch, dbg); break;
    CVB_AR_CASE(1) CVB_AR_CASE(2) CVB_AR_CASE(3) CVB_AR_CASE(4) CVB_AR_CASE(5) CVB_AR_CASE(6) CVB_AR_CASE(7) CVB_AR_CASE(8)
#undef CVB_AR_CASE
  }
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}

extern "C" int cvb_allreduce_wait(const cvb_comm* comm, int n_buckets, uint32_t epoch, void* stream) {
  CommDev c;
  int rc = load_comm(comm, &c);
  if (rc) return rc;
  CVB_REQUIRE(n_buckets >= 0 && n_buckets <= CVB_COMM_MAX_BUCKETS && epoch != 0, CVB_ERR_INVALID_ARG,
              "allreduce_wait: %d buckets (max %d) / epoch %u", n_buckets, CVB_COMM_MAX_BUCKETS, epoch);
  if (n_buckets == 0) return CVB_OK;
  allreduce_wait_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(c, n_buckets, epoch);
  CVB_LAUNCH_CHECK();
  return CVB_OK;
}
