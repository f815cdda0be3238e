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
//   --batch K                frames fused per pass over the volume (default 8)
//   --device N               CUDA device ordinal (default 0)
//   --thresh T W             surface thresholds (reference: 0.2 1.0, demo.cu:194)
//
// All computation runs on the GPU through the C ABI (include/tsdf_fusion.h); host code here only
// reads files and does the 4x4 pose algebra with the reference's rounding.
#include "tsdf_fusion.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
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
  int batch = 8, device = 0;
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
    else if (a == "--thresh") { need(2); tsdf_thresh = atof(argv[i + 1]); weight_thresh = atof(argv[i + 2]); i += 2; }
    else fatal("unknown option " + a);
  }
  if (batch < 1) batch = 1;
  if (batch > TSDF_MAX_BATCH) batch = TSDF_MAX_BATCH;

  // demo.cu:116-117 -- camera intrinsics
  float cam_K[9];
  check(tsdf_load_matrix(cam_K_file.c_str(), 3, 3, cam_K), "LoadMatrixFromFile(cam_K_file)");

  // demo.cu:120-128 -- base frame pose and its inverse (return value ignored like the reference)
  float base2world[16], base2world_inv[16] = {0};
  check(tsdf_load_matrix((frame_prefix(data_path, base_frame_idx) + ".pose.txt").c_str(), 4, 4, base2world),
        "LoadMatrixFromFile(base pose)");
  tsdf_invert_matrix(base2world, base2world_inv);

  // demo.cu:131-153 -- volume initialised directly in HBM
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
  tsdf_volume* vol = nullptr;
  check(tsdf_create(&desc, &vol), "tsdf_create");

  // demo.cu:156-183 -- frame loop; frames are handed to the engine `batch` at a time
  const size_t px = (size_t)im_width * im_height;
  std::vector<uint16_t> depth_mm(px * batch);
  std::vector<float> poses(16 * (size_t)batch);
  int pending = 0;
  auto flush = [&]() {
    if (pending == 0) return;
    check(tsdf_integrate_batch_u16(vol, pending, depth_mm.data(), poses.data()), "Integrate");
    check(tsdf_synchronize(vol), "Integrate");  // the staging vectors are reused
    pending = 0;
  };
  for (int frame_idx = first_frame_idx; frame_idx < first_frame_idx + (int)num_frames; ++frame_idx) {
    const std::string prefix = frame_prefix(data_path, frame_idx);
    const std::string depth_im_file = prefix + ".depth.png";
    if (tsdf_read_depth_png_u16(depth_im_file.c_str(), im_height, im_width, depth_mm.data() + px * pending) != TSDF_OK) {
      std::cout << "Error: depth image file not read!" << std::endl;  // utils.hpp:77
      fatal(tsdf_last_error());
    }
    float cam2world[16];
    check(tsdf_load_matrix((prefix + ".pose.txt").c_str(), 4, 4, cam2world), "LoadMatrixFromFile(pose)");
    tsdf_multiply_matrix(base2world_inv, cam2world, &poses[16 * (size_t)pending]);  // demo.cu:171
    std::cout << "Fusing: " << depth_im_file << std::endl;                          // demo.cu:177
    if (++pending == batch) flush();
  }
  flush();

  // demo.cu:190-194
  std::cout << "Saving surface point cloud (" << ply_path << ")..." << std::endl;
  int64_t n_points = 0;
  check(tsdf_extract_surface(vol, tsdf_thresh, weight_thresh, &n_points), "SaveVoxelGrid2SurfacePointCloud");
  std::vector<float> xyz((size_t)n_points * 3);
  check(tsdf_copy_surface(vol, xyz.data(), n_points), "SaveVoxelGrid2SurfacePointCloud");
  check(tsdf_write_ply(ply_path.c_str(), xyz.data(), n_points), "SaveVoxelGrid2SurfacePointCloud");

  // demo.cu:196-213 (opt-in)
  if (!bin_path.empty()) {
    std::cout << "Saving TSDF voxel grid values to disk (" << bin_path << ")..." << std::endl;
    std::vector<float> tsdf((size_t)dim_x * dim_y * dim_z);
    check(tsdf_download(vol, tsdf.data(), nullptr), "tsdf_download");
    check(tsdf_write_bin(bin_path.c_str(), dim_x, dim_y, dim_z, origin_x, origin_y, origin_z, voxel_size,
                         trunc_margin, tsdf.data()), "tsdf_write_bin");
  }
  tsdf_destroy(vol);
  return 0;
}
