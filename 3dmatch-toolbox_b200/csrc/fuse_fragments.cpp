// fuse_fragments.cpp -- native replacement for the MATLAB fragment loop depth-fusion/fuseSceneFragments.m:36-81
// (BASELINE config 5).
//
//   fuse_fragments <dataPath> <fragmentsPath> [scene ...] [options]
//
// For every scene directory under <dataPath> (or the listed ones), for at most three `seq-*` sequences
// (fuseSceneFragments.m:49), for every window of 50 frames `frameIdx = 0:50:(numFrames-1-50)` (:59): fuse the
// window with base frame = first frame exactly as `./demo camK seqPath frameIdx frameIdx 50 -1.5 -1.5 0.5
// 0.006 0.03` would (:63-64), and write
//   <fragmentsPath>/<scene>/cloud_bin_<N>.ply        (the tsdf.ply the script moves, :66-67)
//   <fragmentsPath>/<scene>/cloud_bin_<N>.info.txt   (scene, sequence, first and last frame index, then the base
//                                                     frame's camera-to-world pose, :70-74)
// The MATLAB loop pays one process start, one 1 GB volume upload and one 1 GB download per fragment; here a few
// engine handles per GPU stay alive, are re-initialised in HBM between fragments, and several fragments are fused by
// ONE launch (tsdf_integrate_multi) so that small volumes fill the GPU together.  Fragments are handed out dynamically
// to --gpus devices (one host thread per device plus PNG decoder threads; fragments are independent, no communication).
//
// Options: --frames N (50)  --voxel V (0.006)  --trunc T (5*V)  --origin X Y Z (-1.5 -1.5 0.5)  --dim N (500)
//          --image W H (640 480)  --batch K (32)  --gpus G (1)  --max-seqs S (3)  --per-launch B (4)
//          --decode-threads T (4 per GPU)  --quiet
#include "tsdf_fusion.h"

#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Options {
  std::string data_path, out_path;
  std::vector<std::string> scenes;
  int frames = 50, dim = 500, im_w = 640, im_h = 480, batch = 32, gpus = 1, max_seqs = 3, per_launch = 4, decode_threads = 4;
  double voxel_d = 0.006, trunc_d = -1.0;  // as the MATLAB script holds them (doubles)
  float voxel = 0.006f, trunc = -1.f, ox = -1.5f, oy = -1.5f, oz = 0.5f;
  bool quiet = false;
};

struct Fragment {
  std::string scene, seq, seq_path, cam_K_file;
  int first_frame, cloud_idx;
};

std::vector<std::string> list_dir(const std::string& path, const std::string& prefix, const std::string& suffix) {
  std::vector<std::string> out;
  if (DIR* d = opendir(path.c_str())) {
    while (dirent* e = readdir(d)) {
      const std::string n = e->d_name;
      if (n == "." || n == "..") continue;
      if (n.size() >= prefix.size() + suffix.size() && n.compare(0, prefix.size(), prefix) == 0 &&
          n.compare(n.size() - suffix.size(), suffix.size(), suffix) == 0)
        out.push_back(n);
    }
    closedir(d);
  }
  std::sort(out.begin(), out.end());  // MATLAB dir() order
  return out;
}

bool is_dir(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

std::string frame_prefix(const std::string& seq_path, int idx) {
  char buf[32];
  snprintf(buf, sizeof buf, "%06d", idx);
  return seq_path + "/frame-" + buf;
}

std::mutex g_print;

// Everything of a fragment that comes from disk: base pose, 50 depth frames (decoded by `threads` threads), poses.
struct Loaded {
  float base2world[16];
  std::vector<uint16_t> depth_mm;
  std::vector<float> poses;
};

bool load_fragment(const Options& o, const Fragment& fr, int threads, Loaded* L, std::string* err) {
  const size_t px = (size_t)o.im_w * o.im_h;
  float base2world_inv[16] = {0};
  if (tsdf_load_matrix((frame_prefix(fr.seq_path, fr.first_frame) + ".pose.txt").c_str(), 4, 4, L->base2world)) {
    *err = tsdf_last_error();
    return false;
  }
  tsdf_invert_matrix(L->base2world, base2world_inv);  // demo.cu:127-128
  L->depth_mm.resize(px * o.frames);
  L->poses.resize(16 * (size_t)o.frames);
  std::atomic<int> next{0};
  std::atomic<bool> ok{true};
  std::mutex err_lock;
  auto work = [&]() {
    for (int i = next++; i < o.frames; i = next++) {  // demo.cu:156-171
      const std::string prefix = frame_prefix(fr.seq_path, fr.first_frame + i);
      float cam2world[16];
      if (tsdf_read_depth_png_u16((prefix + ".depth.png").c_str(), o.im_h, o.im_w, L->depth_mm.data() + px * i) ||
          tsdf_load_matrix((prefix + ".pose.txt").c_str(), 4, 4, cam2world)) {
        std::lock_guard<std::mutex> lock(err_lock);
        *err = tsdf_last_error();
        ok = false;
        return;
      }
      tsdf_multiply_matrix(base2world_inv, cam2world, &L->poses[16 * (size_t)i]);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return ok;
}

// Surface extraction + cloud_bin_N.ply / cloud_bin_N.info.txt of a fused fragment.
bool finish_fragment(const Options& o, const Fragment& fr, const Loaded& L, tsdf_volume* vol, std::string* err) {
  int64_t n_points = 0;
  std::vector<float> xyz;
  if (tsdf_extract_surface(vol, 0.2f, 1.0f, &n_points)) {  // demo.cu:194 thresholds
    *err = tsdf_last_error();
    return false;
  }
  xyz.resize((size_t)n_points * 3);
  const std::string stem = o.out_path + "/" + fr.scene + "/cloud_bin_" + std::to_string(fr.cloud_idx);
  if (tsdf_copy_surface(vol, xyz.data(), n_points) || tsdf_write_ply((stem + ".ply").c_str(), xyz.data(), n_points)) {
    *err = tsdf_last_error();
    return false;
  }
  // fuseSceneFragments.m:70-74: '%s\t %s\t %d\t %d\t\n' then four rows of '%15.8e\t %15.8e\t %15.8e\t %15.8e\t\n'
  FILE* fp = fopen((stem + ".info.txt").c_str(), "w");
  if (!fp) {
    *err = "cannot write " + stem + ".info.txt";
    return false;
  }
  fprintf(fp, "%s\t %s\t %d\t %d\t\n", fr.scene.c_str(), fr.seq.c_str(), fr.first_frame, fr.first_frame + o.frames - 1);
  // the script reads the pose with dlmread (double precision) and prints it; do the same so the digits agree
  double pose[16];
  for (int i = 0; i < 16; ++i) pose[i] = L.base2world[i];
  if (FILE* pf = fopen((frame_prefix(fr.seq_path, fr.first_frame) + ".pose.txt").c_str(), "r")) {
    double tmp[16];
    int got = 0;
    while (got < 16 && fscanf(pf, "%lf", &tmp[got]) == 1) ++got;
    fclose(pf);
    if (got == 16) memcpy(pose, tmp, sizeof pose);
  }
  for (int r = 0; r < 4; ++r)
    fprintf(fp, "%15.8e\t %15.8e\t %15.8e\t %15.8e\t\n", pose[4 * r], pose[4 * r + 1], pose[4 * r + 2],
            pose[4 * r + 3]);
  fclose(fp);
  if (!o.quiet) {
    std::lock_guard<std::mutex> lock(g_print);
    std::cout << "Fused: " << fr.scene << "/" << fr.seq << " frames " << fr.first_frame << ".."
              << fr.first_frame + o.frames - 1 << " -> " << stem << ".ply (" << n_points << " points)" << std::endl;
  }
  return true;
}

}  // namespace

int main(int argc, char* argv[]) {
  Options o;
  std::vector<std::string> pos;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto need = [&](int n) {
      if (i + n >= argc) { std::cerr << "option " << a << " needs " << n << " value(s)" << std::endl; exit(EXIT_FAILURE); }
    };
    if (a == "--frames") { need(1); o.frames = atoi(argv[++i]); }
    else if (a == "--voxel") { need(1); o.voxel_d = atof(argv[++i]); }
    else if (a == "--trunc") { need(1); o.trunc_d = atof(argv[++i]); }
    else if (a == "--origin") { need(3); o.ox = atof(argv[i + 1]); o.oy = atof(argv[i + 2]); o.oz = atof(argv[i + 3]); i += 3; }
    else if (a == "--dim") { need(1); o.dim = atoi(argv[++i]); }
    else if (a == "--image") { need(2); o.im_w = atoi(argv[i + 1]); o.im_h = atoi(argv[i + 2]); i += 2; }
    else if (a == "--batch") { need(1); o.batch = atoi(argv[++i]); }
    else if (a == "--gpus") { need(1); o.gpus = atoi(argv[++i]); }
    else if (a == "--max-seqs") { need(1); o.max_seqs = atoi(argv[++i]); }
    else if (a == "--per-launch") { need(1); o.per_launch = atoi(argv[++i]); }
    else if (a == "--decode-threads") { need(1); o.decode_threads = atoi(argv[++i]); }
    else if (a == "--quiet") o.quiet = true;
    else if (a.compare(0, 2, "--") == 0) { std::cerr << "unknown option " << a << std::endl; return EXIT_FAILURE; }
    else pos.push_back(a);
  }
  if (pos.size() < 2) {
    std::cerr << "usage: fuse_fragments <dataPath> <fragmentsPath> [scene ...] [options]" << std::endl;
    return EXIT_FAILURE;
  }
  o.data_path = pos[0];
  o.out_path = pos[1];
  o.scenes.assign(pos.begin() + 2, pos.end());
  // fuseSceneFragments.m:20-21,63-64: voxelSize and truncMargin = voxelSize*5 are doubles that reach ./demo
  // through sprintf('%f') and atof() -> float; reproduce that path so the fragments are bit-identical
  if (o.trunc_d < 0) o.trunc_d = o.voxel_d * 5;
  {
    char buf[64];
    snprintf(buf, sizeof buf, "%f", o.voxel_d);
    o.voxel = (float)atof(buf);
    snprintf(buf, sizeof buf, "%f", o.trunc_d);
    o.trunc = (float)atof(buf);
  }
  if (o.frames < 1 || o.batch < 1) return EXIT_FAILURE;
  if (o.per_launch < 1) o.per_launch = 1;
  if (o.decode_threads < 1) o.decode_threads = 1;
  if (o.batch > TSDF_MAX_BATCH) o.batch = TSDF_MAX_BATCH;
  if (o.scenes.empty())
    for (const std::string& s : list_dir(o.data_path, "", ""))
      if (is_dir(o.data_path + "/" + s)) o.scenes.push_back(s);

  // enumerate the fragments exactly like the MATLAB loops (fuseSceneFragments.m:36-79)
  std::vector<Fragment> work;
  mkdir(o.out_path.c_str(), 0777);
  for (const std::string& scene : o.scenes) {
    const std::string scene_path = o.data_path + "/" + scene;
    mkdir((o.out_path + "/" + scene).c_str(), 0777);  // :47
    int cloud_idx = 0;
    std::vector<std::string> seqs = list_dir(scene_path, "seq-", "");
    if ((int)seqs.size() > o.max_seqs) seqs.resize(o.max_seqs);  // :49
    for (const std::string& seq : seqs) {
      const std::string seq_path = scene_path + "/" + seq;
      const int n_frames = (int)list_dir(seq_path, "", ".depth.png").size();  // :53
      for (int f = 0; f <= n_frames - 1 - o.frames; f += o.frames)           // :59
        work.push_back({scene, seq, seq_path, scene_path + "/camera-intrinsics.txt", f, cloud_idx++});
    }
  }
  if (!o.quiet) std::cout << work.size() << " fragments, " << o.gpus << " GPU(s)" << std::endl;

  std::atomic<size_t> next{0};
  std::atomic<int> failures{0};
  auto report = [&](const Fragment& fr, const std::string& err) {
    std::lock_guard<std::mutex> lock(g_print);
    std::cerr << "fragment " << fr.scene << "/" << fr.seq << " @" << fr.first_frame << ": " << err << std::endl;
    ++failures;
  };
  auto worker = [&](int device) {
    std::vector<tsdf_volume*> vols;   // per_launch live handles, re-initialised in HBM between fragments
    std::string cur_K;
    for (;;) {
      // a batch: up to per_launch consecutive fragments of the work list
      const size_t w0 = next.fetch_add((size_t)o.per_launch);
      if (w0 >= work.size()) break;
      const size_t w1 = std::min(work.size(), w0 + (size_t)o.per_launch);
      std::vector<const Fragment*> frs;
      std::vector<Loaded> loaded(w1 - w0);
      std::vector<size_t> slot;         // index into `loaded` of the fragments that were read successfully
      for (size_t w = w0; w < w1; ++w) {
        std::string err;
        if (load_fragment(o, work[w], o.decode_threads, &loaded[w - w0], &err)) { frs.push_back(&work[w]); slot.push_back(w - w0); }
        else report(work[w], err);
      }
      // fragments of one launch must share the intrinsics (they are per scene, fuseSceneFragments.m:42): runs of equal K
      size_t i0 = 0;
      while (i0 < frs.size()) {
        size_t i1 = i0 + 1;
        while (i1 < frs.size() && frs[i1]->cam_K_file == frs[i0]->cam_K_file) ++i1;
        std::string err;
        bool ok = true;
        if (frs[i0]->cam_K_file != cur_K || vols.size() < i1 - i0) {
          float cam_K[9];
          if (tsdf_load_matrix(frs[i0]->cam_K_file.c_str(), 3, 3, cam_K)) { err = tsdf_last_error(); ok = false; }
          if (ok && frs[i0]->cam_K_file != cur_K) {
            for (tsdf_volume* v : vols) tsdf_destroy(v);
            vols.clear();
          }
          while (ok && vols.size() < i1 - i0) {
            tsdf_volume_desc d;
            memset(&d, 0, sizeof d);
            d.dim_x = d.dim_y = d.dim_z = o.dim;
            d.origin_x = o.ox; d.origin_y = o.oy; d.origin_z = o.oz;
            d.voxel_size = o.voxel; d.trunc_margin = o.trunc;
            d.im_width = o.im_w; d.im_height = o.im_h;
            memcpy(d.cam_K, cam_K, sizeof cam_K);
            d.device = device; d.max_batch = o.batch; d.cull_level = -1;
            tsdf_volume* v = nullptr;
            if (tsdf_create(&d, &v)) { err = tsdf_last_error(); ok = false; } else vols.push_back(v);
          }
          cur_K = ok ? frs[i0]->cam_K_file : "";
        }
        const int n = (int)(i1 - i0);
        std::vector<const void*> frames((size_t)n);
        std::vector<const float*> poses((size_t)n);
        for (int j = 0; j < n && ok; ++j) {
          frames[j] = loaded[slot[i0 + j]].depth_mm.data();
          poses[j] = loaded[slot[i0 + j]].poses.data();
          if (tsdf_reset(vols[j])) { err = tsdf_last_error(); ok = false; }
        }
        // one launch for the whole run of fragments (demo.cu:156-183 for each of them)
        if (ok && tsdf_integrate_multi(vols.data(), n, o.frames, frames.data(), 1, poses.data())) { err = tsdf_last_error(); ok = false; }
        for (int j = 0; j < n; ++j) {
          if (ok) {
            std::string e2;
            if (!finish_fragment(o, *frs[i0 + j], loaded[slot[i0 + j]], vols[j], &e2)) report(*frs[i0 + j], e2);
          } else {
            report(*frs[i0 + j], err);
          }
        }
        i0 = i1;
      }
    }
    for (tsdf_volume* v : vols) tsdf_destroy(v);
  };
  std::vector<std::thread> threads;
  for (int g = 0; g < o.gpus; ++g) threads.emplace_back(worker, g);
  for (auto& t : threads) t.join();
  if (failures) {
    std::cerr << "FatalError. Program Terminated." << std::endl;
    return EXIT_FAILURE;
  }
  return 0;
}
