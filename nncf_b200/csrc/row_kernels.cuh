// row_kernels.cuh — vectorised (128-bit) versions of the two HBM-bound row kernels of the training step, used when
// the embedding dimension is a multiple of 4 (rows are then 16-byte aligned).  Included by train_step.cu after the
// definitions of GatherArgs / FinalizeArgs.
//
//   gather_rows_vec_kernel    ids -> rows: one LDG.128 per lane per row, FOUR rows in flight per warp (the id -> row
//                             dependency is a double DRAM latency; memory-level parallelism is what hides it)
//   finalize_vec_kernel       corrections / normalise-backward / regulariser, then the sparse SGD update as one
//                             red.global.add.v4.f32 per lane per row (4x fewer L2 atomic operations than scalar)
#pragma once

namespace nncf {

constexpr int kRowsPerWarp = 4;
// The gather runs while the NEXT score kernel's CTAs are already resident (768 threads per SM, prologue overlap): with 4
// rows per warp the gather grid is 2,048 threads per SM at R = 37 and needed 1.6 waves next to them; 8 rows per warp halve
// the grid (1,024 threads per SM): one wave, and twice the loads in flight per warp.
// (a small grid - R = 1 - is faster with 4: 12.7 vs 13.6 us per step: the host picks)

template <int NV, int kGatherRowsPerWarp>   // NV: float4 chunks per lane (1 for dp <= 128, 2 for dp = 256); rows in flight per warp: 4 or 8
__global__ void __launch_bounds__(256)
gather_rows_vec_kernel(GatherArgs a0, GatherArgs a1) {
  pdl_launch_dependents();      // the score kernel may set up its barriers / TMEM while the rows are gathered
  if (a0.tl && threadIdx.x == 0) atomicMin(&a0.tl[0], global_timer_ns());
  const GatherArgs& a = blockIdx.z ? a1 : a0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + warp) * kGatherRowsPerWarp;
  const int r = blockIdx.y;
  // ids of my four rows (one load instruction for the warp).  Link ids are inputs of the step, nothing in the stream
  // writes them: they are fetched BEFORE waiting for the previous kernel, which takes one of the two dependent round
  // trips (id -> row) off the critical path.  (The compact item ids of the group scheme are produced by this step's
  // unique kernel: those wait first.)
  const bool early_ids = a.count_dev == nullptr;
  int64_t myid = 0;
  if (early_ids && row0 < a.rows_pad && lane < kGatherRowsPerWarp && row0 + lane < a.count && a.table)
    myid = a.ids[r * a.ids_stride + row0 + lane];
  pdl_wait();                   // the previous step's update (and this step's tf.unique) must have landed
  if (a0.tl && threadIdx.x == 0) atomicMin(&a0.tl[1], global_timer_ns());
  if (row0 >= a.rows_pad) return;
  const int count = a.count_dev ? a.count_dev[r] : a.count;
  const int nchunk = a.dp / 64;
  if (!early_ids && lane < kGatherRowsPerWarp && row0 + lane < count && a.table) myid = a.ids[r * a.ids_stride + row0 + lane];
  float4 x[kGatherRowsPerWarp][NV];
#pragma unroll
  for (int k = 0; k < kGatherRowsPerWarp; ++k) {
    const int row = row0 + k;
    const int64_t id = __shfl_sync(0xffffffffu, myid, k);
    const float* src = a.shards.n > 1 ? a.shards.p[id % a.shards.n] + (id / a.shards.n) * a.d
                                      : (a.table ? a.table + id * a.d : a.dense_rows + (int64_t)row * a.d);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      x[k][v] = (row < count && c < a.d) ? __ldg(reinterpret_cast<const float4*>(src + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int k = 0; k < kGatherRowsPerWarp; ++k) {
    const int row = row0 + k;
    if (row >= a.rows_pad) break;
    float inv = 1.0f;
    if (a.normalize) {
      float ss = 0.0f;
#pragma unroll
      for (int v = 0; v < NV; ++v) ss += x[k][v].x * x[k][v].x + x[k][v].y * x[k][v].y + x[k][v].z * x[k][v].z + x[k][v].w * x[k][v].w;
      ss = warp_sum(ss);
      inv = (row < count) ? rsqrtf(fmaxf(ss, 1e-12f)) : 1.0f;
#pragma unroll
      for (int v = 0; v < NV; ++v) { x[k][v].x *= inv; x[k][v].y *= inv; x[k][v].z *= inv; x[k][v].w *= inv; }
    }
    const int64_t rowoff = (int64_t)r * a.rows_pad + row;
    uint8_t* blk = a.write_img ? a.img + ((int64_t)r * (a.rows_pad / 128) + (row >> 7)) * nchunk * kSubBytes : nullptr;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      if (c < a.dp) {
        if (a.write_xf) *reinterpret_cast<float4*>(a.Xf + rowoff * a.dp + c) = x[k][v];
        if (a.zero_grad) *reinterpret_cast<float4*>(a.dX + rowoff * a.dp + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.write_img) {
          uint2 pk;
          pk.x = pack_bf16x2(x[k][v].x, x[k][v].y);
          pk.y = pack_bf16x2(x[k][v].z, x[k][v].w);
          *reinterpret_cast<uint2*>(blk + (c >> 6) * kSubBytes + sw128_offset(row & 127, c & 63)) = pk;
        }
      }
    }
    if (lane == 0) {
      a.inv[rowoff] = inv;
      a.corr[rowoff] = 0.0f;
    }
  }
  if (a0.tl && threadIdx.x == 0) atomicMax(&a0.tl[2], global_timer_ns());
}

template <int NV>
__global__ void __launch_bounds__(256)
finalize_vec_kernel(FinalizeArgs a0, FinalizeArgs a1) {
  pdl_launch_dependents();
  pdl_wait();                   // the score kernel's gradient blocks
  const FinalizeArgs& a = blockIdx.z ? a1 : a0;
  finalize_publish_loss(a0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + warp) * kRowsPerWarp;
  const int r = blockIdx.y;
  const int count = a.count_dev ? a.count_dev[r] : a.count;
  if (row0 >= count) return;
  const int64_t base = (int64_t)r * a.rows_pad;
  int64_t myid = 0;
  if (lane < kRowsPerWarp && row0 + lane < count && a.table) myid = a.ids[r * a.ids_stride + row0 + lane];
  float4 g[kRowsPerWarp][NV];
#pragma unroll
  for (int k = 0; k < kRowsPerWarp; ++k) {
    const int row = row0 + k;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      g[k][v] = (row < count && c < a.dp) ? *reinterpret_cast<const float4*>(a.dX + (base + row) * a.dp + c)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int k = 0; k < kRowsPerWarp; ++k) {
    const int row = row0 + k;
    if (row >= count) break;
    float4 xv[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      xv[v] = (a.need_x && c < a.dp) ? *reinterpret_cast<const float4*>(a.Xf + (base + row) * a.dp + c)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (a.pairwise) {
      if (a.scheme == NNCF_SCHEME_NEG_SHARED) {
        const float cs = a.corr_self[base + row];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = 4 * (lane + 32 * v);
          if (c < a.dp) {
            const float4 o = *reinterpret_cast<const float4*>(a.Of + (base + row) * a.dp + c);
            g[k][v].x = fmaf(cs, o.x, g[k][v].x); g[k][v].y = fmaf(cs, o.y, g[k][v].y);
            g[k][v].z = fmaf(cs, o.z, g[k][v].z); g[k][v].w = fmaf(cs, o.w, g[k][v].w);
          }
        }
      } else if (a.side == 0) {
        const float rs = a.corr_self[base + row];
        const int p = a.inverse[base + row];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = 4 * (lane + 32 * v);
          if (c < a.dp) {
            const float4 o = *reinterpret_cast<const float4*>(a.Of + (base + p) * a.dp + c);
            g[k][v].x = fmaf(rs, o.x, g[k][v].x); g[k][v].y = fmaf(rs, o.y, g[k][v].y);
            g[k][v].z = fmaf(rs, o.z, g[k][v].z); g[k][v].w = fmaf(rs, o.w, g[k][v].w);
            red_add_v4(a.dO + (base + p) * a.dp + c, rs * xv[v].x, rs * xv[v].y, rs * xv[v].z, rs * xv[v].w);
          }
        }
      }
    }
    float invn = 1.0f;
    if (a.normalize) {
      invn = a.inv[base + row];
      float dot = 0.0f;
#pragma unroll
      for (int v = 0; v < NV; ++v) dot += g[k][v].x * xv[v].x + g[k][v].y * xv[v].y + g[k][v].z * xv[v].z + g[k][v].w * xv[v].w;
      dot = warp_sum(dot);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        g[k][v].x = (g[k][v].x - xv[v].x * dot) * invn; g[k][v].y = (g[k][v].y - xv[v].y * dot) * invn;
        g[k][v].z = (g[k][v].z - xv[v].z * dot) * invn; g[k][v].w = (g[k][v].w - xv[v].w * dot) * invn;
      }
    }
    if (a.reg_scale != 0.0f) {
      const float s = a.reg_scale / invn;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        g[k][v].x = fmaf(s, xv[v].x, g[k][v].x); g[k][v].y = fmaf(s, xv[v].y, g[k][v].y);
        g[k][v].z = fmaf(s, xv[v].z, g[k][v].z); g[k][v].w = fmaf(s, xv[v].w, g[k][v].w);
      }
    }
    const int64_t id = __shfl_sync(0xffffffffu, myid, k);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (lane + 32 * v);
      if (c >= a.dp) continue;
      if (a.write_back) *reinterpret_cast<float4*>(a.dX + (base + row) * a.dp + c) = g[k][v];
      if (c < a.d) {
        if (a.grad_out && r == 0) *reinterpret_cast<float4*>(a.grad_out + (int64_t)row * a.d + c) = g[k][v];
        if (a.optimizer == NNCF_OPT_SGD && a.table)
          red_add_v4((a.shards.n > 1 ? a.shards.p[id % a.shards.n] + (id / a.shards.n) * a.d : a.table + id * a.d) + c, -a.lr * g[k][v].x, -a.lr * g[k][v].y, -a.lr * g[k][v].z, -a.lr * g[k][v].w);
      }
    }
  }
}

}  // namespace nncf
