// tsdf_sharded.cu -- one TSDF volume cut into z-slabs over the GPUs of a box, behind the C ABI (include/tsdf_fusion.h,
// tsdf_sharded_*).  SURVEY.md 8e / BASELINE north_star (4): voxels never interact and frames are read-only, so the
// volume is cut into contiguous z-slabs (z is the slowest index, depth-fusion/demo.cu:61: a slab is one contiguous
// byte range of both arrays); depth frames are uploaded ONCE to the first slab's device and broadcast to the others
// with ncclBroadcast over NVLink, every slab fuses all frames independently with the global z index, and slabs are
// gathered only for output: surface points by ordered ncclSend / ncclRecv to the first device (rank order == ascending
// z == ascending linear index, so the concatenation is exactly the reference's tsdf.ply payload, utils.hpp:36-53),
// tsdf.bin by slab-ordered pwrite.
//
// One host thread drives all devices (every call below only enqueues work); NCCL is loaded at run time
// (dlopen("libnccl.so.2")), so the library has no link-time dependency on it and single-GPU users never touch it.
// The same device may hold several slabs (used by the tests on one-GPU boxes); slabs that share a device with the
// root exchange data with plain device copies instead of NCCL.
#include "tsdf_engine_internal.h"

#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) {
      auto sym = [&](const char* n) { return dlsym(api.lib, n); };
      api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
      api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
      api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
      api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
      api.Send = (decltype(api.Send))sym("ncclSend");
      api.Recv = (decltype(api.Recv))sym("ncclRecv");
      api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
      api.ok = api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Broadcast && api.Send &&
               api.Recv && api.GetErrorString;
    }
  }
  return api;
}

#define NK(expr)                                                                                                   \
  do {                                                                                                             \
    ncclResult_t r__ = (expr);                                                                                     \
    if (r__ != ncclSuccess)                                                                                        \
      return set_err(TSDF_ERR_CUDA, "%s:%d: %s -> NCCL: %s", __FILE__, __LINE__, #expr, nccl().GetErrorString(r__)); \
  } while (0)

constexpr int kGroupFrames = 128;  // frames per broadcast (double-buffered against the fusion of the previous group)

}  // namespace

struct tsdf_sharded {
  tsdf_volume_desc d;
  int n = 0;
  std::vector<int> dev;
  std::vector<int64_t> cuts;
  std::vector<tsdf_volume*> vol;
  bool use_nccl = false;
  std::vector<ncclComm_t> comm;           // one per slab (NCCL rank == slab index)
  std::vector<cudaStream_t> bstream;      // broadcast / gather stream per slab
  std::vector<char*> stage[2];            // per slab: two staging buffers of kGroupFrames frames
  std::vector<cudaEvent_t> ev_bcast[2];   // the group has arrived in stage[b][i]
  std::vector<cudaEvent_t> ev_fused[2];   // slab i has fused the group it read from stage[b][i]
  bool fused_valid[2] = {false, false};
  size_t stage_bytes = 0;
  int64_t group = 0;
  float* xyz = nullptr;                   // gathered surface points on dev[0]
  int64_t xyz_cap = 0, n_points = -1;
};

namespace {

int ensure_stage(tsdf_sharded* s, size_t bytes) {
  if (s->stage_bytes >= bytes) return TSDF_OK;
  for (int b = 0; b < 2; ++b)
    for (int i = 0; i < s->n; ++i) {
      CK(cudaSetDevice(s->dev[i]));
      if (s->stage[b][i]) CK(cudaFree(s->stage[b][i]));
      s->stage[b][i] = nullptr;
      CK(cudaMalloc(&s->stage[b][i], bytes));
    }
  s->stage_bytes = bytes;
  return TSDF_OK;
}

}  // namespace

extern "C" {

int tsdf_sharded_create(const tsdf_volume_desc* desc, int32_t n_slabs, const int32_t* devices, const int64_t* z_cuts,
                        tsdf_sharded** out) {
  if (!desc || !out || n_slabs < 1) return set_err(TSDF_ERR_ARG, "bad argument");
  *out = nullptr;
  if (n_slabs > desc->dim_z) return set_err(TSDF_ERR_ARG, "more slabs than z-planes");
  int n_dev = 0;
  CK(cudaGetDeviceCount(&n_dev));
  tsdf_sharded* s = new (std::nothrow) tsdf_sharded();
  if (!s) return set_err(TSDF_ERR_ARG, "out of host memory");
  s->d = *desc;
  s->n = n_slabs;
  auto fail = [&](int code) { tsdf_sharded_destroy(s); return code; };
  bool distinct = true;
  for (int i = 0; i < n_slabs; ++i) {
    const int dv = devices ? devices[i] : i;
    if (dv < 0 || dv >= n_dev) return fail(set_err(TSDF_ERR_ARG, "device %d of slab %d does not exist (%d devices)", dv, i, n_dev));
    for (int j = 0; j < i; ++j) distinct = distinct && s->dev[j] != dv;
    s->dev.push_back(dv);
  }
  s->cuts.resize((size_t)n_slabs + 1);
  for (int i = 0; i <= n_slabs; ++i) s->cuts[i] = z_cuts ? z_cuts[i] : (desc->dim_z * i) / n_slabs;
  if (s->cuts[0] != 0 || s->cuts[n_slabs] != desc->dim_z) return fail(set_err(TSDF_ERR_ARG, "z_cuts must run from 0 to dim_z"));
  for (int i = 0; i < n_slabs; ++i)
    if (s->cuts[i + 1] <= s->cuts[i]) return fail(set_err(TSDF_ERR_ARG, "z_cuts must be strictly ascending"));
  s->vol.assign((size_t)n_slabs, nullptr);
  s->bstream.assign((size_t)n_slabs, nullptr);
  for (int b = 0; b < 2; ++b) {
    s->stage[b].assign((size_t)n_slabs, nullptr);
    s->ev_bcast[b].assign((size_t)n_slabs, nullptr);
    s->ev_fused[b].assign((size_t)n_slabs, nullptr);
  }
  for (int i = 0; i < n_slabs; ++i) {
    tsdf_volume_desc d = *desc;
    d.device = s->dev[i];
    d.z_begin = s->cuts[i];
    d.z_end = s->cuts[i + 1];
    d.stream = nullptr;
    d.ext_tsdf = d.ext_weight = nullptr;
    if (int rc = tsdf_create(&d, &s->vol[i])) return fail(rc);
    cudaError_t e = cudaSetDevice(s->dev[i]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->bstream[i], cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
      e = cudaEventCreateWithFlags(&s->ev_bcast[b][i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_fused[b][i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return fail(set_err(TSDF_ERR_CUDA, "slab %d on device %d: %s", i, s->dev[i], cudaGetErrorString(e)));
  }
  if (n_slabs > 1 && distinct) {
    if (!nccl().ok) return fail(set_err(TSDF_ERR_CUDA, "libnccl.so.2 could not be loaded (needed for %d devices)", n_slabs));
    s->comm.assign((size_t)n_slabs, nullptr);
    ncclResult_t r = nccl().CommInitAll(s->comm.data(), n_slabs, s->dev.data());
    if (r != ncclSuccess) {
      s->comm.clear();
      return fail(set_err(TSDF_ERR_CUDA, "ncclCommInitAll: %s", nccl().GetErrorString(r)));
    }
    s->use_nccl = true;
  }
  *out = s;
  return TSDF_OK;
}

void tsdf_sharded_destroy(tsdf_sharded* s) {
  if (!s) return;
  for (int i = 0; i < (int)s->vol.size(); ++i) {
    if (i < (int)s->dev.size()) cudaSetDevice(s->dev[i]);
    if (i < (int)s->bstream.size() && s->bstream[i]) cudaStreamSynchronize(s->bstream[i]);
  }
  for (ncclComm_t c : s->comm)
    if (c) nccl().CommDestroy(c);
  for (int i = 0; i < (int)s->vol.size(); ++i) {
    tsdf_destroy(s->vol[i]);
    if (i < (int)s->dev.size()) cudaSetDevice(s->dev[i]);
    for (int b = 0; b < 2; ++b) {
      if (i < (int)s->stage[b].size()) cudaFree(s->stage[b][i]);
      if (i < (int)s->ev_bcast[b].size() && s->ev_bcast[b][i]) cudaEventDestroy(s->ev_bcast[b][i]);
      if (i < (int)s->ev_fused[b].size() && s->ev_fused[b][i]) cudaEventDestroy(s->ev_fused[b][i]);
    }
    if (i < (int)s->bstream.size() && s->bstream[i]) cudaStreamDestroy(s->bstream[i]);
  }
  if (s->xyz && !s->dev.empty()) { cudaSetDevice(s->dev[0]); cudaFree(s->xyz); }
  cudaGetLastError();
  delete s;
}

int tsdf_sharded_slab(tsdf_sharded* s, int32_t i, tsdf_volume** vol, int64_t* z_begin, int64_t* z_end, int32_t* device) {
  if (!s || i < 0 || i >= s->n) return set_err(TSDF_ERR_ARG, "slab index out of range");
  if (vol) *vol = s->vol[i];
  if (z_begin) *z_begin = s->cuts[i];
  if (z_end) *z_end = s->cuts[i + 1];
  if (device) *device = s->dev[i];
  return TSDF_OK;
}

int tsdf_sharded_reset(tsdf_sharded* s) {
  if (!s) return set_err(TSDF_ERR_ARG, "handle is NULL");
  for (tsdf_volume* v : s->vol)
    if (int rc = tsdf_reset(v)) return rc;
  s->n_points = -1;
  return TSDF_OK;
}

int tsdf_sharded_synchronize(tsdf_sharded* s) {
  if (!s) return set_err(TSDF_ERR_ARG, "handle is NULL");
  for (int i = 0; i < s->n; ++i) {
    CK(cudaSetDevice(s->dev[i]));
    CK(cudaStreamSynchronize(s->bstream[i]));
    if (int rc = tsdf_synchronize(s->vol[i])) return rc;
  }
  return TSDF_OK;
}

int tsdf_sharded_integrate_batch(tsdf_sharded* s, int32_t n_frames, const void* frames, int32_t frames_are_u16,
                                 const float* cam2base) {
  if (!s || !frames || !cam2base) return set_err(TSDF_ERR_ARG, "NULL argument");
  if (n_frames < 0) return set_err(TSDF_ERR_ARG, "n_frames < 0");
  const size_t px = (size_t)s->d.im_width * s->d.im_height;
  const size_t elem = frames_are_u16 ? sizeof(uint16_t) : sizeof(float);
  if (s->n == 1) {  // nothing to broadcast
    return frames_are_u16 ? tsdf_integrate_batch_u16(s->vol[0], n_frames, (const uint16_t*)frames, cam2base)
                          : tsdf_integrate_batch(s->vol[0], n_frames, (const float*)frames, cam2base);
  }
  if (int rc = ensure_stage(s, (size_t)kGroupFrames * px * sizeof(float))) return rc;
  for (int g0 = 0; g0 < n_frames; g0 += kGroupFrames) {
    const int ng = (n_frames - g0 < kGroupFrames) ? n_frames - g0 : kGroupFrames;
    const size_t bytes = (size_t)ng * px * elem;
    const int b = (int)(s->group++ & 1);
    // the staging buffers of this parity were last read by the fusion of the group before the previous one
    if (s->fused_valid[b])
      for (int i = 0; i < s->n; ++i) {
        CK(cudaSetDevice(s->dev[i]));
        CK(cudaStreamWaitEvent(s->bstream[i], s->ev_fused[b][i], 0));
        if (i != 0) CK(cudaStreamWaitEvent(s->bstream[0], s->ev_fused[b][i], 0));  // the root overwrites what peers still pull
      }
    // one upload (or device copy) to the first slab's device ...
    CK(cudaSetDevice(s->dev[0]));
    CK(cudaMemcpyAsync(s->stage[b][0], (const char*)frames + (size_t)g0 * px * elem, bytes, cudaMemcpyDefault, s->bstream[0]));
    s->vol[0]->stats.h2d_bytes += (int64_t)bytes;
    CK(cudaEventRecord(s->ev_bcast[b][0], s->bstream[0]));
    // ... then the broadcast to every other slab's device
    if (s->use_nccl) {
      NK(nccl().GroupStart());
      for (int i = 0; i < s->n; ++i) {
        ncclResult_t r = nccl().Broadcast(s->stage[b][0], s->stage[b][i], bytes, ncclUint8, 0, s->comm[i], s->bstream[i]);
        if (r != ncclSuccess) { nccl().GroupEnd(); NK(r); }
      }
      NK(nccl().GroupEnd());
    } else {
      for (int i = 1; i < s->n; ++i) {
        CK(cudaSetDevice(s->dev[i]));
        CK(cudaStreamWaitEvent(s->bstream[i], s->ev_bcast[b][0], 0));
        if (s->stage[b][i] != s->stage[b][0])
          CK(cudaMemcpyAsync(s->stage[b][i], s->stage[b][0], bytes, cudaMemcpyDefault, s->bstream[i]));
      }
    }
    for (int i = 0; i < s->n; ++i) {
      CK(cudaSetDevice(s->dev[i]));
      CK(cudaEventRecord(s->ev_bcast[b][i], s->bstream[i]));
      // the slab's engine stream waits for the group, fuses it (asynchronously), and marks the buffer free
      CK(cudaStreamWaitEvent(s->vol[i]->stream, s->ev_bcast[b][i], 0));
      const int rc = frames_are_u16
                         ? tsdf_integrate_batch_u16(s->vol[i], ng, (const uint16_t*)s->stage[b][i], cam2base + 16 * (size_t)g0)
                         : tsdf_integrate_batch(s->vol[i], ng, (const float*)s->stage[b][i], cam2base + 16 * (size_t)g0);
      if (rc) return rc;
      CK(cudaSetDevice(s->dev[i]));
      CK(cudaEventRecord(s->ev_fused[b][i], s->vol[i]->stream));
    }
    s->fused_valid[b] = true;
  }
  s->n_points = -1;
  return TSDF_OK;
}

int tsdf_sharded_extract_surface(tsdf_sharded* s, float tsdf_thresh, float weight_thresh, int64_t* n_points) {
  if (!s) return set_err(TSDF_ERR_ARG, "handle is NULL");
  std::vector<int64_t> count((size_t)s->n, 0);
  int64_t total = 0;
  for (int i = 0; i < s->n; ++i) {
    if (int rc = tsdf_extract_surface(s->vol[i], tsdf_thresh, weight_thresh, &count[i])) return rc;
    total += count[i];
  }
  CK(cudaSetDevice(s->dev[0]));
  if (total > s->xyz_cap) {
    if (s->xyz) CK(cudaFree(s->xyz));
    s->xyz = nullptr;
    s->xyz_cap = 0;
    CK(cudaMalloc(&s->xyz, (size_t)total * 3 * sizeof(float)));
    s->xyz_cap = total;
  }
  // ordered gather: slab order == ascending z == ascending linear voxel index (utils.hpp:36-53)
  int64_t off = 0;
  for (int i = 0; i < s->n; ++i) {
    float* src = nullptr;
    int64_t n = 0;
    if (int rc = tsdf_surface_device_ptr(s->vol[i], &src, &n)) return rc;  // (waits for the slab's extraction kernels)
    if (n > 0) {
      float* dst = s->xyz + 3 * off;
      if (s->use_nccl && i != 0) {
        NK(nccl().GroupStart());
        ncclResult_t r1 = nccl().Send(src, (size_t)n * 3, ncclFloat, 0, s->comm[i], s->bstream[i]);
        ncclResult_t r2 = nccl().Recv(dst, (size_t)n * 3, ncclFloat, i, s->comm[0], s->bstream[0]);
        NK(nccl().GroupEnd());
        NK(r1);
        NK(r2);
      } else {
        CK(cudaSetDevice(s->dev[0]));
        CK(cudaMemcpyAsync(dst, src, (size_t)n * 3 * sizeof(float), cudaMemcpyDefault, s->bstream[0]));
      }
    }
    off += n;
  }
  for (int i = 0; i < s->n; ++i) {
    CK(cudaSetDevice(s->dev[i]));
    CK(cudaStreamSynchronize(s->bstream[i]));
  }
  s->n_points = total;
  if (n_points) *n_points = total;
  return TSDF_OK;
}

int tsdf_sharded_copy_surface(tsdf_sharded* s, float* xyz_host, int64_t capacity_points) {
  if (!s) return set_err(TSDF_ERR_ARG, "handle is NULL");
  if (s->n_points < 0) return set_err(TSDF_ERR_STATE, "tsdf_sharded_extract_surface has not been called since the last update");
  if (capacity_points < s->n_points) return set_err(TSDF_ERR_ARG, "capacity %lld < %lld points", (long long)capacity_points, (long long)s->n_points);
  if (s->n_points == 0) return TSDF_OK;
  if (!xyz_host) return set_err(TSDF_ERR_ARG, "xyz_host is NULL");
  CK(cudaSetDevice(s->dev[0]));
  CK(cudaMemcpy(xyz_host, s->xyz, (size_t)s->n_points * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  s->vol[0]->stats.d2h_bytes += s->n_points * 12;
  return TSDF_OK;
}

int tsdf_sharded_write_bin(tsdf_sharded* s, const char* path) {
  if (!s || !path) return set_err(TSDF_ERR_ARG, "NULL argument");
  if (int rc = tsdf_sharded_synchronize(s)) return rc;
  const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) return set_err(TSDF_ERR_IO, "cannot open %s for writing", path);
  // demo.cu:196-213: 8-float header, then dim^3 TSDF floats in linear order (x fastest): slab i owns the byte range
  // [32 + 4 * z_cuts[i] * plane, 32 + 4 * z_cuts[i + 1] * plane)
  const float header[8] = {(float)s->d.dim_x, (float)s->d.dim_y, (float)s->d.dim_z, s->d.origin_x, s->d.origin_y,
                           s->d.origin_z, s->d.voxel_size, s->d.trunc_margin};
  int rc = TSDF_OK;
  if (pwrite(fd, header, sizeof header, 0) != (ssize_t)sizeof header) rc = set_err(TSDF_ERR_IO, "short write to %s", path);
  const int64_t plane = s->d.dim_x * s->d.dim_y;
  const int64_t piece = ((64ll << 20) / (plane * 4) > 0) ? (64ll << 20) / (plane * 4) : 1;  // planes per copy (~64 MB)
  float* host = nullptr;
  if (!rc && cudaHostAlloc((void**)&host, (size_t)piece * plane * sizeof(float), cudaHostAllocDefault) != cudaSuccess)
    rc = set_err(TSDF_ERR_CUDA, "cudaHostAlloc of the tsdf.bin staging buffer failed");
  for (int i = 0; i < s->n && !rc; ++i) {
    float* t_dev = nullptr;
    tsdf_device_ptrs(s->vol[i], &t_dev, nullptr);
    cudaSetDevice(s->dev[i]);
    for (int64_t z = s->cuts[i]; z < s->cuts[i + 1] && !rc; z += piece) {
      const int64_t nzp = (s->cuts[i + 1] - z < piece) ? s->cuts[i + 1] - z : piece;
      const size_t bytes = (size_t)nzp * plane * sizeof(float);
      if (cudaMemcpy(host, t_dev + (z - s->cuts[i]) * plane, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
        rc = set_err(TSDF_ERR_CUDA, "download of slab %d failed: %s", i, cudaGetErrorString(cudaGetLastError()));
        break;
      }
      s->vol[i]->stats.d2h_bytes += (int64_t)bytes;
      size_t done = 0;
      while (done < bytes) {
        const ssize_t w = pwrite(fd, (const char*)host + done, bytes - done, (off_t)(32 + z * plane * 4 + (int64_t)done));
        if (w <= 0) { rc = set_err(TSDF_ERR_IO, "short write to %s", path); break; }
        done += (size_t)w;
      }
    }
  }
  if (host) cudaFreeHost(host);
  if (close(fd) != 0 && !rc) rc = set_err(TSDF_ERR_IO, "close of %s failed", path);
  return rc;
}

}  // extern "C"
