// demo.cpp -- drop-in replacement for the reference CLI `depth-fusion/demo` (depth-fusion/demo.cu:72-216).
//
//   demo [cam_K_file data_path base_frame_idx first_frame_idx num_frames
//         origin_x origin_y origin_z voxel_size trunc_margin] [options]
//
// argv[1..10] are the reference's positional arguments (demo.cu:102-113; all ten or none, as
// depth-fusion/fuseSceneFragments.m:63-64 passes them); with no arguments the reference's built-in
// defaults apply (demo.cu:75-99).  Inputs: <cam_K_file> (3x3 text), <data_path>/frame-%06d.pose.txt
// (4x4 text, camera->world), <data_path>/frame-%06d.depth.png (16-bit, millimetres).  Outputs:
// tsdf.ply in the working directory (utils.hpp:26-53), stdout lines `Fusing: <depth file>` per frame
// and `Saving surface point cloud (tsdf.ply)...` (demo.cu:177,191).
//
// Options (only AFTER the ten positionals, so existing callers are unaffected):
//   --dims X Y Z | --dim N   grid size (reference: 500 500 500, demo.cu:97-99)
//   --image W H              depth image size (reference: 640 480, demo.cu:87-88)
//   --bin [path]             also write the TSDF volume as tsdf.bin (demo.cu:196-213)
//   --ply path               surface point cloud path (default tsdf.ply)
//   --batch K                frames per pass over the volume (default 32)
//   --device N               CUDA device ordinal (default 0)
//   --gpus N                 cut the volume into N z-slabs of equal work over devices 0 .. N-1 (frames broadcast with
//                            NCCL, slabs gathered only for output; tsdf.ply / tsdf.bin are byte-identical to --gpus 1)
//   --threads T              PNG decoder threads (default: hardware concurrency, at most 16)
//   --timing                 phase times (JSON, one line) on stderr: CUDA context + volume init, waiting for the decoder,
//                            fusion tail, extraction, file output
//   --thresh T W             surface thresholds (reference: 0.2 1.0, demo.cu:194)
//
// All computation runs on the GPU through the C ABI (include/tsdf_fusion.h); host code here only
// reads files and does the 4x4 pose algebra with the reference's rounding.
#include "tsdf_fusion.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

namespace {

[[noreturn]] void fatal(const std::string& what) {
  // utils.hpp:240-253: message on stderr, then EXIT_FAILURE
  std::cerr << what << std::endl;
  std::cerr << "FatalError. Program Terminated." << std::endl;
  exit(EXIT_FAILURE);
}

void check(int rc, const char* what) {
  if (rc == TSDF_OK) return;
  if (rc == TSDF_ERR_CUDA) fatal(std::string("CUDA failure in ") + what + ": " + tsdf_last_error());
  fatal(std::string(what) + ": " + tsdf_last_error());
}

std::string frame_prefix(const std::string& data_path, int idx) {
  char buf[32];
  snprintf(buf, sizeof buf, "%06d", idx);  // demo.cu:119-121: setw(6) setfill('0')
  return data_path + "/frame-" + buf;
}

}  // namespace

int main(int argc, char* argv[]) {
  // demo.cu:75-81
  std::string cam_K_file = "../data/sample/depth-fusion-demo/camera-intrinsics.txt";
  std::string data_path = "../data/sample/depth-fusion-demo/rgbd-frames";
  int base_frame_idx = 8;
  int first_frame_idx = 8;
  float num_frames = 50;
  // demo.cu:87-99
  int im_width = 640, im_height = 480;
  float origin_x = -1.5f, origin_y = -1.5f, origin_z = 0.5f;
  float voxel_size = 0.006f;
  float trunc_margin = voxel_size * 5;
  long long dim_x = 500, dim_y = 500, dim_z = 500;
  // extensions
  std::string ply_path = "tsdf.ply", bin_path;
  int batch = 32, device = 0, gpus = 1, threads = 0;
  bool timing = false;
  const auto t_start = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point t) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count();
  };
  float tsdf_thresh = 0.2f, weight_thresh = 1.0f;  // demo.cu:194

  int next = 1;
  if (argc > 1 && strncmp(argv[1], "--", 2) != 0) {
    if (argc < 11) fatal("usage: demo cam_K_file data_path base_frame_idx first_frame_idx num_frames "
                         "origin_x origin_y origin_z voxel_size trunc_margin [options]");
    // demo.cu:102-113
    cam_K_file = argv[1];
    data_path = argv[2];
    base_frame_idx = atoi(argv[3]);
    first_frame_idx = atoi(argv[4]);
    num_frames = atof(argv[5]);
    origin_x = atof(argv[6]);
    origin_y = atof(argv[7]);
    origin_z = atof(argv[8]);
    voxel_size = atof(argv[9]);
    trunc_margin = atof(argv[10]);
    next = 11;
  }
  for (int i = next; i < argc; ++i) {
    const std::string a = argv[i];
    auto need = [&](int n) { if (i + n >= argc) fatal("option " + a + " needs " + std::to_string(n) + " value(s)"); };
    if (a == "--dims") { need(3); dim_x = atoll(argv[i + 1]); dim_y = atoll(argv[i + 2]); dim_z = atoll(argv[i + 3]); i += 3; }
    else if (a == "--dim") { need(1); dim_x = dim_y = dim_z = atoll(argv[i + 1]); i += 1; }
    else if (a == "--image") { need(2); im_width = atoi(argv[i + 1]); im_height = atoi(argv[i + 2]); i += 2; }
    else if (a == "--bin") { if (i + 1 < argc && strncmp(argv[i + 1], "--", 2) != 0) bin_path = argv[++i]; else bin_path = "tsdf.bin"; }
    else if (a == "--ply") { need(1); ply_path = argv[++i]; }
    else if (a == "--batch") { need(1); batch = atoi(argv[++i]); }
    else if (a == "--device") { need(1); device = atoi(argv[++i]); }
    else if (a == "--gpus") { need(1); gpus = atoi(argv[++i]); }
    else if (a == "--threads") { need(1); threads = atoi(argv[++i]); }
    else if (a == "--timing") timing = true;
    else if (a == "--thresh") { need(2); tsdf_thresh = atof(argv[i + 1]); weight_thresh = atof(argv[i + 2]); i += 2; }
    else fatal("unknown option " + a);
  }
  if (batch < 1) batch = 1;
  if (batch > TSDF_MAX_BATCH) batch = TSDF_MAX_BATCH;
  if (gpus < 1) gpus = 1;
  if (threads < 1) {
    threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 4;
    if (threads > 16) threads = 16;
  }
  const int n_frames = (int)num_frames > 0 ? (int)num_frames : 0;

  // demo.cu:116-117 -- camera intrinsics
  float cam_K[9];
  check(tsdf_load_matrix(cam_K_file.c_str(), 3, 3, cam_K), "LoadMatrixFromFile(cam_K_file)");

  // demo.cu:120-128 -- base frame pose and its inverse (return value ignored like the reference)
  float base2world[16], base2world_inv[16] = {0};
  check(tsdf_load_matrix((frame_prefix(data_path, base_frame_idx) + ".pose.txt").c_str(), 4, 4, base2world),
        "LoadMatrixFromFile(base pose)");
  tsdf_invert_matrix(base2world, base2world_inv);

  tsdf_volume_desc desc;
  memset(&desc, 0, sizeof desc);
  desc.dim_x = dim_x; desc.dim_y = dim_y; desc.dim_z = dim_z;
  desc.origin_x = origin_x; desc.origin_y = origin_y; desc.origin_z = origin_z;
  desc.voxel_size = voxel_size; desc.trunc_margin = trunc_margin;
  desc.im_width = im_width; desc.im_height = im_height;
  memcpy(desc.cam_K, cam_K, sizeof cam_K);
  desc.device = device;
  desc.max_batch = batch;
  desc.cull_level = -1;

  // demo.cu:156-171 -- the frames are decoded by a pool of threads (PNG inflate is the slowest host step) into one
  // buffer that lives until the end of the run, so nothing has to wait for the GPU before it is reused; the main
  // thread hands every completed group of frames to the engine, which fuses it asynchronously while the pool decodes
  // the next ones
  const size_t px = (size_t)im_width * im_height;
  std::vector<uint16_t> depth_mm(px * (size_t)n_frames);
  std::vector<float> poses(16 * (size_t)n_frames);
  std::vector<std::atomic<int>> done((size_t)n_frames);
  for (auto& d : done) d.store(0);
  std::vector<std::string> errors((size_t)n_frames);
  std::atomic<int> next_frame{0};
  auto decode = [&]() {
    for (int i = next_frame++; i < n_frames; i = next_frame++) {
      const std::string prefix = frame_prefix(data_path, first_frame_idx + i);
      float cam2world[16];
      if (tsdf_read_depth_png_u16((prefix + ".depth.png").c_str(), im_height, im_width, depth_mm.data() + px * (size_t)i) != TSDF_OK) {
        errors[i] = std::string("Error: depth image file not read!\n") + tsdf_last_error();   // utils.hpp:77
        done[i].store(2);
        continue;
      }
      if (tsdf_load_matrix((prefix + ".pose.txt").c_str(), 4, 4, cam2world) != TSDF_OK) {
        errors[i] = std::string("LoadMatrixFromFile(pose): ") + tsdf_last_error();
        done[i].store(2);
        continue;
      }
      tsdf_multiply_matrix(base2world_inv, cam2world, &poses[16 * (size_t)i]);  // demo.cu:171
      done[i].store(1);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < threads && t < n_frames; ++t) pool.emplace_back(decode);
  auto wait_frames = [&](int lo, int hi) {   // frames [lo, hi) decoded; prints the reference's progress lines in order
    for (int i = lo; i < hi; ++i) {
      while (done[i].load() == 0) std::this_thread::sleep_for(std::chrono::microseconds(200));
      if (done[i].load() == 2) {
        for (auto& t : pool) t.join();
        if (errors[i].rfind("Error:", 0) == 0) {
          std::cout << "Error: depth image file not read!" << std::endl;
          fatal(errors[i].substr(errors[i].find('\n') + 1));
        }
        fatal(errors[i]);
      }
      std::cout << "Fusing: " << frame_prefix(data_path, first_frame_idx + i) << ".depth.png" << std::endl;  // demo.cu:177
    }
  };

  tsdf_volume* vol = nullptr;
  tsdf_sharded* shard = nullptr;
  double t_init = 0, t_decode_wait = 0;
  if (gpus == 1) {
    // demo.cu:131-153 -- volume initialised directly in HBM
    const auto t0 = std::chrono::steady_clock::now();
    check(tsdf_create(&desc, &vol), "tsdf_create");
    t_init = since(t0);
    const int group = 32;
    for (int f0 = 0; f0 < n_frames; f0 += group) {
      const int n = (n_frames - f0 < group) ? n_frames - f0 : group;
      const auto tw = std::chrono::steady_clock::now();
      wait_frames(f0, f0 + n);
      t_decode_wait += since(tw);
      check(tsdf_integrate_batch_u16(vol, n, depth_mm.data() + px * (size_t)f0, poses.data() + 16 * (size_t)f0), "Integrate");
    }
  } else {
    // equal-work z-slabs need the whole trajectory: decode first, then plan, then fuse
    const auto tw = std::chrono::steady_clock::now();
    wait_frames(0, n_frames);
    t_decode_wait = since(tw);
    const auto t0 = std::chrono::steady_clock::now();
    const int n_groups = (int)((dim_z + 7) / 8);
    std::vector<uint64_t> cost((size_t)n_groups, 1);
    std::vector<int64_t> cuts((size_t)gpus + 1);
    if (n_frames > 0)
      check(tsdf_plan_z_cost(&desc, n_frames, depth_mm.data(), 1, poses.data(), n_frames > 128 ? n_frames / 128 : 1,
                             cost.data(), n_groups), "tsdf_plan_z_cost");
    check(tsdf_split_slabs(cost.data(), n_groups, dim_z, gpus, 0.05, cuts.data()), "tsdf_split_slabs");
    check(tsdf_sharded_create(&desc, gpus, nullptr, cuts.data(), &shard), "tsdf_sharded_create");
    t_init = since(t0);
    check(tsdf_sharded_integrate_batch(shard, n_frames, depth_mm.data(), 1, poses.data()), "Integrate");
  }
  const auto t_fuse0 = std::chrono::steady_clock::now();
  if (timing) check(vol ? tsdf_synchronize(vol) : tsdf_sharded_synchronize(shard), "synchronize");
  const double t_fuse_tail = since(t_fuse0);
  for (auto& t : pool) t.join();

  // demo.cu:190-194
  std::cout << "Saving surface point cloud (" << ply_path << ")..." << std::endl;
  int64_t n_points = 0;
  std::vector<float> xyz;
  if (vol) {
    check(tsdf_extract_surface(vol, tsdf_thresh, weight_thresh, &n_points), "SaveVoxelGrid2SurfacePointCloud");
    xyz.resize((size_t)n_points * 3);
    check(tsdf_copy_surface(vol, xyz.data(), n_points), "SaveVoxelGrid2SurfacePointCloud");
  } else {
    check(tsdf_sharded_extract_surface(shard, tsdf_thresh, weight_thresh, &n_points), "SaveVoxelGrid2SurfacePointCloud");
    xyz.resize((size_t)n_points * 3);
    check(tsdf_sharded_copy_surface(shard, xyz.data(), n_points), "SaveVoxelGrid2SurfacePointCloud");
  }
  const double t_extract = since(t_fuse0) - t_fuse_tail;
  const auto t_w0 = std::chrono::steady_clock::now();
  check(tsdf_write_ply(ply_path.c_str(), xyz.data(), n_points), "SaveVoxelGrid2SurfacePointCloud");

  // demo.cu:196-213 (opt-in)
  if (!bin_path.empty()) {
    std::cout << "Saving TSDF voxel grid values to disk (" << bin_path << ")..." << std::endl;
    if (vol) {
      std::vector<float> tsdf((size_t)dim_x * dim_y * dim_z);
      check(tsdf_download(vol, tsdf.data(), nullptr), "tsdf_download");
      check(tsdf_write_bin(bin_path.c_str(), dim_x, dim_y, dim_z, origin_x, origin_y, origin_z, voxel_size,
                           trunc_margin, tsdf.data()), "tsdf_write_bin");
    } else {
      check(tsdf_sharded_write_bin(shard, bin_path.c_str()), "tsdf_write_bin");
    }
  }
  const double t_write = since(t_w0);
  tsdf_destroy(vol);
  tsdf_sharded_destroy(shard);
  if (timing)
    fprintf(stderr, "{\"frames\": %d, \"decoder_threads\": %d, \"gpus\": %d, \"init_s\": %.4f, \"decode_wait_s\": %.4f, "
            "\"fusion_tail_s\": %.4f, \"extract_s\": %.4f, \"write_s\": %.4f, \"total_s\": %.4f}\n",
            n_frames, threads, gpus, t_init, t_decode_wait, t_fuse_tail, t_extract, t_write, since(t_start));
  return 0;
}
