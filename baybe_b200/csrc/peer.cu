// peer.cu -- the global arg-max of a row-sharded candidate set without a host-issued collective.
//
// Every rank's scoring kernel leaves its shard's packed (score, lowest global index) key in local memory.  One
// warp then (k_allreduce_best)
//   1. publishes: lane r does atomicMax.sys on rank r's key slot and, after a system fence, bumps rank r's arrival
//      counter -- plain NVLink peer stores through pointers the caller obtained with CUDA IPC (the slots live in one
//      small buffer per rank);
//   2. waits until its OWN counter shows `world` arrivals for this epoch (bounded spin: ~11 s, then status = 1);
//   3. copies the reduced key out and re-arms its slot for the epoch after next.
// Slots and counters are double-buffered by epoch parity; counters only grow (arrivals of use k of a parity end at
// world * k), so nothing is ever reset while a peer may still write it:
//   * my slot[p] is re-armed after MY wait of epoch e (p = e & 1) -- by then all ranks have written it for e;
//   * a peer writes my slot[p] again in epoch e + 2, which it can only reach after I contributed to e + 1, i.e.
//     after my kernel of epoch e (re-arm included) in stream order.
// Replaces the per-round ncclAllReduce(MAX, int64) + host launch of round 1 (1 -> 8 GPU weak-scaling efficiency 0.55):
// the step becomes [scoring kernel][this kernel] on one stream, no host in between, no NCCL proxy.
// Reference semantics replaced: torch.argmax over the concatenated candidate set inside
// botorch.optim.optimize_acqf_discrete (baybe/recommenders/pure/bayesian/botorch/discrete.py:124-126).
#include "common.cuh"

namespace bb {

__global__ void __launch_bounds__(32) k_allreduce_best(bb_peer_group g, const long long* __restrict__ local_key,
                                                       uint32_t epoch, long long* __restrict__ out_key,
                                                       int* __restrict__ status) {
  const int lane = threadIdx.x;
  const uint32_t par = epoch & 1u;
  const uint32_t target = (uint32_t)g.world * ((epoch >> 1) + 1u);
  const long long mine = *local_key;
  if (lane < g.world) {
    long long* slot = reinterpret_cast<long long*>(g.d_key[lane]) + par;
    if (mine != kEmptyKey) atomicMax_system(slot, mine);
    __threadfence_system();
    atomicAdd_system(g.d_count[lane] + par, 1u);
  }
  __syncwarp();
  if (lane == 0) {
    volatile uint32_t* cnt = reinterpret_cast<volatile uint32_t*>(g.d_count[g.rank] + par);
    const long long t0 = clock64();
    bool ok = true;
    while ((int32_t)(*cnt - target) < 0) {
      if (clock64() - t0 > (20ll << 30)) {  // ~11 s at 2 GHz: a peer never arrived -- report, do not hang the GPU
                                            // (ranks that enter their first reduction seconds apart are normal)
        ok = false;
        break;
      }
      __nanosleep(200);
    }
    __threadfence_system();
    long long* slot = reinterpret_cast<long long*>(g.d_key[g.rank]) + par;
    const long long v = *reinterpret_cast<volatile long long*>(slot);
    *out_key = ok ? v : mine;
    *status = ok ? 0 : 1;
    *slot = kEmptyKey;  // re-arm for epoch + 2
    __threadfence_system();
  }
}

__global__ void k_peer_init(long long* keys, uint32_t* counts) {
  if (threadIdx.x < 2) {
    keys[threadIdx.x] = kEmptyKey;
    counts[threadIdx.x] = 0u;
  }
}

}  // namespace bb

using namespace bb;

extern "C" int bb_peer_slots_init(int64_t* d_keys, uint32_t* d_counts, void* stream) {
  BB_CHECK_ARG(d_keys && d_counts, "bb_peer_slots_init: null buffer");
  k_peer_init<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(d_keys), d_counts);
  BB_LAUNCH_CHECK();
  return BB_OK;
}

extern "C" int bb_allreduce_best(const bb_peer_group* g, const int64_t* d_local_key, uint32_t epoch,
                                 int64_t* d_out_key, int32_t* d_status, void* stream) {
  BB_CHECK_ARG(g && d_local_key && d_out_key && d_status, "bb_allreduce_best: null argument");
  BB_CHECK_ARG(g->world >= 1 && g->world <= BB_MAX_PEERS && g->rank >= 0 && g->rank < g->world,
               "bb_allreduce_best: rank %d / world %d outside [0,%d]", g->rank, g->world, BB_MAX_PEERS);
  for (int r = 0; r < g->world; ++r)
    BB_CHECK_ARG(g->d_key[r] && g->d_count[r], "bb_allreduce_best: slot pointers of rank %d are null", r);
  k_allreduce_best<<<1, 32, 0, (cudaStream_t)stream>>>(*g, reinterpret_cast<const long long*>(d_local_key), epoch,
                                                       reinterpret_cast<long long*>(d_out_key), d_status);
  BB_LAUNCH_CHECK();
  return BB_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Level-coded candidate rows -> fp32 rows.  A discrete search space has a handful of distinct values per column
// (the parameter's value list, baybe/searchspace/discrete.py:529-536), so a row is d small integers: 4 bits per
// column when every column has <= 16 levels, else 8.  Decoding is exact -- the table holds the float32 comp-rep
// values themselves -- and turns the PCIe-bound end-to-end pass (80 B per candidate at d = 20) into 10 or 20 B.
// Pure HBM streaming: N*(d*bits/8) bytes in, N*d*4 bytes out.
// ------------------------------------------------------------------------------------------------------------
namespace bb {

template <int BITS>
__global__ void __launch_bounds__(256) k_decode_codes(const uint8_t* __restrict__ codes, int64_t N, int d,
                                                      int64_t ld_bytes, const float* __restrict__ table, int table_ld,
                                                      float* __restrict__ out, int64_t ldo) {
  extern __shared__ float tab_s[];  // [d][1 << BITS]
  constexpr int L = 1 << BITS;
  for (int e = threadIdx.x; e < d * L; e += blockDim.x) {
    const int j = e / L, l = e - j * L;
    tab_s[e] = l < table_ld ? __ldg(table + (size_t)j * table_ld + l) : 0.f;
  }
  __syncthreads();
  const int64_t total = N * (int64_t)d;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / d;
    const int j = (int)(e - r * d);
    int code;
    if constexpr (BITS == 8) code = codes[r * ld_bytes + j];
    else code = (codes[r * ld_bytes + (j >> 1)] >> ((j & 1) * 4)) & 15;
    out[r * ldo + j] = tab_s[j * L + code];
  }
}

}  // namespace bb

extern "C" int bb_decode_codes(const uint8_t* d_codes, int32_t bits, int64_t N, int32_t d, int64_t ld_bytes,
                               const float* d_table, int32_t table_ld, float* d_out, int64_t ldo, void* stream) {
  BB_CHECK_ARG(bits == 4 || bits == 8, "bb_decode_codes: bits must be 4 or 8 (got %d)", bits);
  BB_CHECK_ARG(N >= 0 && d >= 1 && d <= 1024, "bb_decode_codes: bad shape N=%lld d=%d", (long long)N, d);
  BB_CHECK_ARG(ld_bytes >= (bits == 8 ? d : (d + 1) / 2) && ldo >= d, "bb_decode_codes: leading dimension too small");
  BB_CHECK_ARG(table_ld >= 1 && table_ld <= (1 << bits), "bb_decode_codes: table_ld=%d outside [1,%d]", table_ld,
               1 << bits);
  if (N == 0) return BB_OK;
  BB_CHECK_ARG(d_codes && d_table && d_out, "bb_decode_codes: null buffer");
  const size_t smem = (size_t)d * (1u << bits) * sizeof(float);
  BB_CHECK_SUPPORTED(smem <= 48 * 1024, "bb_decode_codes: value table of %zu bytes exceeds 48 KB", smem);
  const int64_t total = N * (int64_t)d;
  int grid = (int)((total + 256 * 8 - 1) / (256 * 8));
  if (grid > kSMs * 8) grid = kSMs * 8;
  if (grid < 1) grid = 1;
  if (bits == 8)
    k_decode_codes<8><<<grid, 256, smem, (cudaStream_t)stream>>>(d_codes, N, d, ld_bytes, d_table, table_ld, d_out, ldo);
  else
    k_decode_codes<4><<<grid, 256, smem, (cudaStream_t)stream>>>(d_codes, N, d, ld_bytes, d_table, table_ld, d_out, ldo);
  BB_LAUNCH_CHECK();
  return BB_OK;
}
