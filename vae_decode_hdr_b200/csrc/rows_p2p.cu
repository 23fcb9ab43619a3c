// Device-driven exchanges of the row-tiled decode (BASELINE config C4: one image split into row slabs over the GPUs of
// one NVSwitch box).  Every rank's workspace is a library-owned allocation that the other ranks map over
// CUDA IPC (hdrvae_peer_alloc / hdrvae_peer_open), so an exchange is two small kernels on the rank's own stream and no
// host round trip, no NCCL call:
//
//   push : (1) tell every rank "I have reached exchange s" (all my kernels that read what exchange s-1 delivered are
//              done: stream order) and wait until every rank has said so — only then may anybody's buffers be overwritten;
//          (2) store this rank's contribution straight into the peers' workspaces over NVLink: the first / last interior
//              row of the conv output(s) into the neighbours' halo rows, the 64 GroupNorm sums and the HDR statistics block
//              into this rank's slot of every rank's table, the attention q|k / v rows into every rank's gather buffers;
//          (3) after a system-scope fence, the last block to finish tells every rank "my data for exchange s has landed".
//   wait : wait until every rank's data flag shows s, then fold the table slots IN RANK ORDER (the sums are therefore
//          bit-identical on every rank and independent of any library's reduction order) into the place the next kernel
//          reads, and advance the exchange counter, which lives in device memory (the program needs no host state).
//
// The two-phase handshake makes the transfer race-free whatever the relative speed of the ranks; the tables are
// double-buffered by the parity of s because a rank may already run push(s+1) while another still folds the table of s.
// Never run two ranks of this transport as separate processes on ONE GPU (their kernels wait on one another); single-GPU
// tests use the host-driven emulation in sharding.py instead.
#include "engine.cuh"

namespace hdrvae {

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// bounded spin: a missing rank must surface as a trapped kernel, not as a hung GPU (~4 s at 2 GHz)
__device__ __forceinline__ void spin_until_ge(const unsigned int* flag, unsigned int s, const char* what) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - s) < 0) {
    if (clock64() - t0 > 8000000000ll) {
      printf("hdrvae rows: timed out waiting for the %s flag (want %u, have %u)\n", what, s, ld_acquire_sys(flag));
      __trap();
    }
    __nanosleep(64);
  }
}

__global__ void __launch_bounds__(256) rows_push_kernel(const RowsPushArgs a) {
  RowsMailbox* mail = reinterpret_cast<RowsMailbox*>(a.ws + a.off_mail);
  const unsigned int s = ld_acquire_sys(&mail->seq) + 1;
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && tid < a.world)
    st_release_sys(&reinterpret_cast<RowsMailbox*>(a.peers[tid] + a.off_mail)->arrived[a.rank], s);
  if (tid < a.world) spin_until_ge(&mail->arrived[tid], s, "arrival");
  __syncthreads();
  const long long gtid = (long long)blockIdx.x * blockDim.x + tid, gsize = (long long)gridDim.x * blockDim.x;
  for (int i = 0; i < a.n_seg; ++i) {
    const RowsSegment sg = a.seg[i];
    unsigned long long dst_off = sg.dst_off;
    if (sg.kind == 1) dst_off = a.off_mail + offsetof(RowsMailbox, sums) + ((size_t)(s & 1) * kRowsMaxRanks + a.rank) * sizeof(mail->sums[0][0]);
    if (sg.kind == 2) dst_off = a.off_mail + offsetof(RowsMailbox, raw) + ((size_t)(s & 1) * kRowsMaxRanks + a.rank) * sizeof(mail->raw[0][0]);
    const uint4* src = reinterpret_cast<const uint4*>(a.ws + sg.src_off);
    uint4* dst = reinterpret_cast<uint4*>(a.peers[sg.peer] + dst_off);
    const long long n16 = sg.bytes >> 4;
    for (long long k = gtid; k < n16; k += gsize) dst[k] = src[k];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  if (tid == 0) {
    const unsigned int t = atomicAdd(&mail->ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1 : 0;
    if (s_last) mail->ticket = 0;
  }
  __syncthreads();
  if (s_last && tid < a.world) {
    __threadfence_system();
    st_release_sys(&reinterpret_cast<RowsMailbox*>(a.peers[tid] + a.off_mail)->data[a.rank], s);
  }
}

__global__ void __launch_bounds__(64) rows_wait_kernel(uint8_t* ws, unsigned long long off_mail, int world,
                                                       unsigned long long allreduce_off, int allreduce_count,
                                                       unsigned long long raw_off, int has_raw) {
  RowsMailbox* mail = reinterpret_cast<RowsMailbox*>(ws + off_mail);
  const unsigned int s = ld_acquire_sys(&mail->seq) + 1;
  const int tid = threadIdx.x;
  if (tid < world) spin_until_ge(&mail->data[tid], s, "data");
  __syncthreads();
  __threadfence_system();
  const int par = s & 1;
  if (allreduce_count > 0) {
    double* out = reinterpret_cast<double*>(ws + allreduce_off);
    for (int k = tid; k < allreduce_count; k += blockDim.x) {
      double v = mail->sums[par][0][k];
      for (int r = 1; r < world; ++r) v += mail->sums[par][r][k];
      out[k] = v;
    }
  }
  if (has_raw) {
    hdrvae_raw_stats* out = reinterpret_cast<hdrvae_raw_stats*>(ws + raw_off);
    const RowsRawSlot* slot = mail->raw[par];
    if (tid < HDRVAE_RAW_NMIN) {
      float v = slot[0].s.vmin[tid];
      for (int r = 1; r < world; ++r) v = fminf(v, slot[r].s.vmin[tid]);
      out->vmin[tid] = v;
    } else if (tid < HDRVAE_RAW_NMIN + HDRVAE_RAW_NMAX) {
      const int k = tid - HDRVAE_RAW_NMIN;
      float v = slot[0].s.vmax[k];
      for (int r = 1; r < world; ++r) v = fmaxf(v, slot[r].s.vmax[k]);
      out->vmax[k] = v;
    } else if (tid < HDRVAE_RAW_NMIN + HDRVAE_RAW_NMAX + HDRVAE_RAW_NSUM) {
      const int k = tid - HDRVAE_RAW_NMIN - HDRVAE_RAW_NMAX;
      double v = slot[0].s.vsum[k];
      for (int r = 1; r < world; ++r) v += slot[r].s.vsum[k];
      out->vsum[k] = v;
    }
  }
  __syncthreads();
  if (tid == 0) st_release_sys(&mail->seq, s);
}

int launch_rows_push(const RowsPushArgs& a, int blocks, cudaStream_t s) {
  rows_push_kernel<<<blocks, 256, 0, s>>>(a);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_rows_wait(uint8_t* ws, size_t off_mail, int world, size_t allreduce_off, int allreduce_count, size_t raw_off,
                     bool has_raw, cudaStream_t s) {
  rows_wait_kernel<<<1, 64, 0, s>>>(ws, off_mail, world, allreduce_off, allreduce_count, raw_off, has_raw ? 1 : 0);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace hdrvae
