// Block-diagonal CSR-by-destination for every graph of the batch in one launch.
// Replaces the per-graph Python slicing of src/aligner/sg_aligner.py:86-104 and PyG's
// remove_self_loops + add_self_loops (torch_geometric 2.2.0 utils/loop.py) inside GATConv.
#include "common.cuh"

namespace sga {
namespace {

constexpr int NT = 256;

// One CTA per graph.  Counting sort by destination in shared memory; rows are then put in input
// order (deterministic summation order downstream) and the self loop is appended last, which is
// where PyG's add_self_loops puts it.
__global__ void __launch_bounds__(NT)
csr_build_kernel(const int64_t* __restrict__ edges, const int32_t* __restrict__ node_off,
                 const int64_t* __restrict__ edge_off, int32_t* __restrict__ row_beg,
                 int32_t* __restrict__ row_cnt, int32_t* __restrict__ col, int max_nodes) {
  extern __shared__ int sm_i[];
  int* cnt = sm_i;               // [max_nodes]
  int* off = sm_i + max_nodes;   // [max_nodes]
  const int g = blockIdx.x;
  const int n0 = node_off[g];
  const int n = node_off[g + 1] - n0;
  const int64_t e0 = edge_off[g];
  const int e = (int)(edge_off[g + 1] - e0);
  const int slot0 = (int)(e0 + n0);
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += NT) cnt[i] = 0;
  __syncthreads();
  for (int k = tid; k < e; k += NT) {
    int s = (int)edges[(e0 + k) * 2], d = (int)edges[(e0 + k) * 2 + 1];
    if (s != d && (unsigned)s < (unsigned)n && (unsigned)d < (unsigned)n) atomicAdd(&cnt[d], 1);
  }
  __syncthreads();
  if (tid < 32) {   // exclusive scan of (cnt+1) by one warp
    int carry = 0;
    for (int b = 0; b < n; b += 32) {
      int i = b + tid;
      int v = (i < n) ? cnt[i] + 1 : 0;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (tid >= o) x += y;
      }
      if (i < n) off[i] = carry + x - v;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += NT) {
    row_beg[n0 + i] = slot0 + off[i];
    row_cnt[n0 + i] = cnt[i] + 1;
    cnt[i] = 0;   // reuse as fill cursor
  }
  __syncthreads();
  for (int k = tid; k < e; k += NT) {
    int s = (int)edges[(e0 + k) * 2], d = (int)edges[(e0 + k) * 2 + 1];
    if (s != d && (unsigned)s < (unsigned)n && (unsigned)d < (unsigned)n) {
      int slot = atomicAdd(&cnt[d], 1);
      col[slot0 + off[d] + slot] = k;   // edge id for now
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += NT) {
    int32_t* r = col + slot0 + off[i];
    const int c = cnt[i];
    for (int a = 1; a < c; ++a) {   // insertion sort of edge ids (rows are short)
      int v = r[a], b = a - 1;
      while (b >= 0 && r[b] > v) { r[b + 1] = r[b]; --b; }
      r[b + 1] = v;
    }
    for (int a = 0; a < c; ++a) r[a] = n0 + (int)edges[(e0 + r[a]) * 2];
    r[c] = n0 + i;   // self loop
  }
}

}  // namespace
}  // namespace sga

extern "C" int sga_csr_build(const int64_t* edges, const int32_t* node_off, const int64_t* edge_off,
                             int G, int max_graph_nodes, int32_t* row_beg, int32_t* row_cnt,
                             int32_t* col, void* stream) {
  if (G <= 0) return SGA_OK;
  // the largest graph bounds the shared-memory scratch (scene graphs: tens to hundreds of nodes)
  SGA_REQUIRE(max_graph_nodes > 0 && max_graph_nodes <= 24 * 1024, "sga_csr_build: max_graph_nodes=%d out of range (1..24576)", max_graph_nodes);
  const int max_nodes = (max_graph_nodes + 31) & ~31;
  size_t smem = 2 * (size_t)max_nodes * sizeof(int);
  static size_t attr_smem = 48 * 1024;
  if (smem > attr_smem) {
    SGA_CUDA(cudaFuncSetAttribute(sga::csr_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  sga::csr_build_kernel<<<G, sga::NT, smem, (cudaStream_t)stream>>>(edges, node_off, edge_off, row_beg, row_cnt, col, max_nodes);
  SGA_LAUNCH_CHECK();
  return SGA_OK;
}
