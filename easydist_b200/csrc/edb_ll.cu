// Low-latency small-message collectives of libedb.so ("LL" protocol).
//
// The epoch-flag protocol of edb_reshard.cu costs two flag round trips and three system-scope
// fences per op (~18 us for a 1 KiB all-reduce at n=2, NCCL: 11 us).  Small reshard edges are
// common in auto-SPMD plans (LayerNorm / bias gradients: 16 x 4 KiB all-reduces per step in the
// reference's own GPT example), so messages up to `ll_max_bytes` per rank use
// packets instead: every 4-byte payload word travels with the 4-byte op number in ONE 8-byte
// store into the receiver's scratch buffer (peer HBM over NVLink).  The receiver polls its own
// memory until the number matches — data and flag arrive together, so there is no fence, no
// separate flag, and nobody reads peer memory.  Buffers are double-buffered by op parity.
// Semantics are those of the callables in easydist/torch/passes/sharding.py:94-152; reductions
// accumulate in rank order in fp32 exactly like k_reduce, so results are bit-identical to the
// large-message path and to the oracle.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "edb_internal.cuh"

namespace edb {

enum { LL_ALL_REDUCE = 0, LL_ALL_GATHER = 1, LL_REDUCE_SCATTER = 2 };

struct LLDesc {
  uint64_t* flags;            // my flag block of the group (F_LLSEQ, F_CNT_LL)
  uint2* ll_peer[kMaxGroup];  // every member's LL region (ll_peer[me] is local)
  int n, me, mode;
  const uint32_t* src;
  uint32_t* dst;
  int64_t words;        // 4-byte words of the local input
  int64_t outer, cw;    // AG / RS geometry: local input is [outer, (n,) cw] words
  int dtype, redop;
  float scale;
  uint64_t timeout_ns;
};

__device__ __forceinline__ void st_pkt(uint2* p, uint32_t data, uint32_t flag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(flag) : "memory");
}
__device__ __forceinline__ uint2 ld_pkt(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t wait_pkt(const uint2* p, uint32_t flag, uint64_t timeout_ns,
                                             uint64_t* err) {
  uint2 v = ld_pkt(p);
  if (v.y == flag) return v.x;
  const uint64_t t0 = globaltimer_ns();
  while (true) {
    v = ld_pkt(p);
    if (v.y == flag) return v.x;
    if (globaltimer_ns() - t0 > timeout_ns) fatal_timeout(err, flag, 3);  // never returns
  }
}

// 4-byte word <-> accumulators
template <int DT> struct Word;
template <> struct Word<EDB_F32> {
  typedef float acc;
  static constexpr int N = 1;
  static __device__ __forceinline__ void unpack(uint32_t w, float* a) { a[0] = __uint_as_float(w); }
  static __device__ __forceinline__ uint32_t pack(const float* a) { return __float_as_uint(a[0]); }
};
template <> struct Word<EDB_I32> {
  typedef int32_t acc;
  static constexpr int N = 1;
  static __device__ __forceinline__ void unpack(uint32_t w, int32_t* a) { a[0] = (int32_t)w; }
  static __device__ __forceinline__ uint32_t pack(const int32_t* a) { return (uint32_t)a[0]; }
};
template <> struct Word<EDB_BF16> {
  typedef float acc;
  static constexpr int N = 2;
  static __device__ __forceinline__ void unpack(uint32_t w, float* a) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
    a[0] = f.x;
    a[1] = f.y;
  }
  static __device__ __forceinline__ uint32_t pack(const float* a) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a[0], a[1]);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <> struct Word<EDB_F16> {
  typedef float acc;
  static constexpr int N = 2;
  static __device__ __forceinline__ void unpack(uint32_t w, float* a) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    a[0] = f.x;
    a[1] = f.y;
  }
  static __device__ __forceinline__ uint32_t pack(const float* a) {
    __half2 h = __floats2half2_rn(a[0], a[1]);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

template <int OP, typename A> __device__ __forceinline__ A ll_combine(A a, A b) {
  if (OP == EDB_MAX) return a > b ? a : b;
  if (OP == EDB_MIN) return a < b ? a : b;
  return a + b;
}

template <int DT, int OP>
__global__ void __launch_bounds__(256) k_ll(const __grid_constant__ LLDesc d) {
  typedef Word<DT> W;
  typedef typename W::acc A;
  __shared__ uint64_t s_q;
  if (threadIdx.x == 0) s_q = ld_relaxed_gpu(d.flags + F_LLSEQ) + 1;
  __syncthreads();
  const uint64_t q = s_q;
  const uint32_t flag = (uint32_t)q;
  const size_t par_off = (size_t)(q & 1) * kMaxGroup * kLLCapacity;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const int n = d.n, me = d.me;

  // ---- push: my words into the receivers' buffer [parity][me] ------------------------------
  if (d.mode == LL_REDUCE_SCATTER) {
    // word (o, p, c) of [outer, n, cw] goes to member p only, at index o*cw + c
    for (int64_t i = tid; i < d.words; i += nthr) {
      const int64_t o = i / (n * d.cw), r = i - o * n * d.cw;
      const int p = (int)(r / d.cw);
      const int64_t c = r - (int64_t)p * d.cw;
      if (p != me) st_pkt(d.ll_peer[p] + par_off + (size_t)me * kLLCapacity + o * d.cw + c, d.src[i], flag);
    }
  } else {
    for (int64_t i = tid; i < d.words; i += nthr) {
      const uint32_t w = d.src[i];
      for (int k = 1; k < n; ++k) {
        const int p = (me + k) % n;
        st_pkt(d.ll_peer[p] + par_off + (size_t)me * kLLCapacity + i, w, flag);
      }
    }
  }

  // ---- receive ----------------------------------------------------------------------------------
  const uint2* mine = d.ll_peer[me] + par_off;
  uint64_t* err = d.flags + F_ERR;
  if (d.mode == LL_ALL_GATHER) {
    // out is [outer, n, cw] words: slot p of every row comes from member p
    for (int64_t i = tid; i < d.words; i += nthr) {
      const int64_t o = i / d.cw, c = i - o * d.cw;
      for (int p = 0; p < n; ++p) {
        const uint32_t w = (p == me) ? d.src[i]
                                     : wait_pkt(mine + (size_t)p * kLLCapacity + i, flag, d.timeout_ns, err);
        d.dst[(o * n + p) * d.cw + c] = w;
      }
    }
  } else {
    const int64_t out_words = (d.mode == LL_REDUCE_SCATTER) ? d.outer * d.cw : d.words;
    for (int64_t i = tid; i < out_words; i += nthr) {
      int64_t own_idx = i;
      if (d.mode == LL_REDUCE_SCATTER) {
        const int64_t o = i / d.cw, c = i - o * d.cw;
        own_idx = (o * n + me) * d.cw + c;
      }
      A acc[W::N], v[W::N];
      for (int p = 0; p < n; ++p) {
        const uint32_t w = (p == me) ? d.src[own_idx]
                                     : wait_pkt(mine + (size_t)p * kLLCapacity + i, flag, d.timeout_ns, err);
        W::unpack(w, v);
        if (p == 0) {
#pragma unroll
          for (int e = 0; e < W::N; ++e) acc[e] = v[e];
        } else {
#pragma unroll
          for (int e = 0; e < W::N; ++e) acc[e] = ll_combine<OP, A>(acc[e], v[e]);
        }
      }
      if (d.scale != 1.0f) {
#pragma unroll
        for (int e = 0; e < W::N; ++e) acc[e] = (A)((float)acc[e] * d.scale);
      }
      d.dst[i] = W::pack(acc);
    }
  }

  // ---- advance the op number (last CTA) -----------------------------------------------------------
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long prev =
        atomicAdd(reinterpret_cast<unsigned long long*>(d.flags + F_CNT_LL), 1ULL);
    if (prev == (unsigned long long)gridDim.x - 1) {
      d.flags[F_CNT_LL] = 0;
      __threadfence();
      st_release_gpu(d.flags + F_LLSEQ, q);
    }
  }
}

template <int DT> static void launch_ll(const LLDesc& d, int grid, cudaStream_t st) {
  switch (d.redop) {
    case EDB_MAX: k_ll<DT, EDB_MAX><<<grid, 256, 0, st>>>(d); break;
    case EDB_MIN: k_ll<DT, EDB_MIN><<<grid, 256, 0, st>>>(d); break;
    default: k_ll<DT, EDB_SUM><<<grid, 256, 0, st>>>(d); break;
  }
}

// Returns EDB_OK after launching, or -1 when the op is not eligible for the LL path.
int ll_try(int gid, int mode, void* dst, const void* src, int64_t in_bytes, int64_t outer,
           int64_t row_bytes, int dtype, int redop, float scale, cudaStream_t st) {
  Runtime& r = rt();
  const Group& g = r.groups[gid];
  if (g.n <= 1 || g.slot >= kLLGroups || r.ll_max_bytes <= 0) return -1;
  if (in_bytes <= 0 || in_bytes > r.ll_max_bytes || (in_bytes & 3) || (row_bytes & 3)) return -1;
  // Eligibility must only depend on rank-invariant quantities (sizes, dtype, group slot, options
  // that the host sets identically everywhere): a rank that picked the flag protocol while its
  // peers use packets would never rendezvous.  Alignment is therefore a contract of the callers
  // (every entry point requires 16-byte aligned bases), not a reason to switch protocols.
  if (((uintptr_t)src | (uintptr_t)dst) & 3)
    return set_error(EDB_E_INVALID, "collective operands must be at least 4-byte aligned");
  if (mode != LL_ALL_GATHER &&
      !(dtype == EDB_F32 || dtype == EDB_BF16 || dtype == EDB_F16 || dtype == EDB_I32))
    return -1;
  if (redop == EDB_AVG && dtype == EDB_I32) return -1;
  const int64_t words = in_bytes / 4;
  const int64_t per_src = (mode == LL_REDUCE_SCATTER) ? words / g.n : words;
  if (per_src > (int64_t)kLLCapacity) return -1;
  LLDesc d;
  memset(&d, 0, sizeof(d));
  d.flags = flag_block(r.heap, g.slot);
  for (int p = 0; p < g.n; ++p)
    d.ll_peer[p] = reinterpret_cast<uint2*>(ll_region(r.peer_heap[g.ranks[p]], g.slot));
  d.n = g.n;
  d.me = g.me;
  d.mode = mode;
  d.src = static_cast<const uint32_t*>(src);
  d.dst = static_cast<uint32_t*>(dst);
  d.words = words;
  d.outer = outer;
  d.cw = row_bytes / 4;
  d.dtype = (mode == LL_ALL_GATHER) ? EDB_I32 : dtype;
  d.redop = (redop == EDB_AVG) ? EDB_SUM : redop;
  d.scale = scale;
  d.timeout_ns = (uint64_t)r.spin_timeout_ms * 1000000ull;
  int grid = (int)((words + 1023) / 1024);
  if (grid > 64) grid = 64;
  if (grid < 1) grid = 1;
  switch (d.dtype) {
    case EDB_F32: launch_ll<EDB_F32>(d, grid, st); break;
    case EDB_BF16: launch_ll<EDB_BF16>(d, grid, st); break;
    case EDB_F16: launch_ll<EDB_F16>(d, grid, st); break;
    default: launch_ll<EDB_I32>(d, grid, st); break;
  }
  count_launch();
  return cuda_check(cudaGetLastError(), "k_ll launch");
}

}  // namespace edb
