// ingest.cuh — wire / disk formats at the seam ("next" row): sensor_msgs/PointCloud2 blobs and binary PCD bodies are
// uploaded AS THEY ARE and unpacked into PointXYZI records on the device, instead of the host-side pcl::fromROSMsg /
// pcl::io::loadPCDFile conversions at /root/reference/apps/scan_matching_odometry_nodelet.cpp:118-119,
// apps/hdl_graph_slam_nodelet.cpp:153-154 and src/hdl_graph_slam/keyframe.cpp:57,141.
//   k_unpack_points   <- pcl::fromROSMsg / the PCD reader's field mapping: x, y, z (FLOAT32) and intensity (FLOAT32 / UINT8 /
//                        UINT16 / UINT32 ...) picked from arbitrary byte offsets of a point_step-long record, optional byte swap
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "engine.cuh"

namespace b2r {

struct UnpackArgs {
  const unsigned char* data;
  long long n;
  unsigned int point_step, off_x, off_y, off_z, off_i, type_i, bigendian;
  float* out;  // 8 floats per point: x y z 1 | intensity 0 0 0  (pcl::PointXYZI)
};

__device__ __forceinline__ unsigned int load_u32(const unsigned char* p, bool swap) {
  unsigned int v = (unsigned int)p[0] | ((unsigned int)p[1] << 8) | ((unsigned int)p[2] << 16) | ((unsigned int)p[3] << 24);  // unaligned-safe
  return swap ? __byte_perm(v, 0, 0x0123) : v;
}

__global__ void k_unpack_points(const __grid_constant__ UnpackArgs A) {
  const bool swap = A.bigendian != 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned char* r = A.data + (size_t)i * A.point_step;
    float* o = A.out + (size_t)i * 8;
    o[0] = __uint_as_float(load_u32(r + A.off_x, swap));
    o[1] = __uint_as_float(load_u32(r + A.off_y, swap));
    o[2] = __uint_as_float(load_u32(r + A.off_z, swap));
    o[3] = 1.0f;
    float inten = 0.f;
    if (A.off_i != 0xffffffffu) {
      const unsigned char* q = r + A.off_i;
      switch (A.type_i) {  // sensor_msgs/PointField datatypes
        case 1: inten = (float)(signed char)q[0]; break;                                   // INT8
        case 2: inten = (float)q[0]; break;                                                // UINT8
        case 3: { unsigned int v = swap ? ((unsigned int)q[0] << 8 | q[1]) : ((unsigned int)q[1] << 8 | q[0]); inten = (float)(short)v; break; }  // INT16
        case 4: { unsigned int v = swap ? ((unsigned int)q[0] << 8 | q[1]) : ((unsigned int)q[1] << 8 | q[0]); inten = (float)v; break; }         // UINT16
        case 5: inten = (float)(int)load_u32(q, swap); break;                              // INT32
        case 6: inten = (float)load_u32(q, swap); break;                                   // UINT32
        case 8: { unsigned long long lo = load_u32(q, false), hi = load_u32(q + 4, false);                                                    // FLOAT64
                  unsigned long long v = swap ? ((unsigned long long)__byte_perm((unsigned int)lo, 0, 0x0123) << 32 | __byte_perm((unsigned int)hi, 0, 0x0123)) : (hi << 32 | lo);
                  inten = (float)__longlong_as_double((long long)v); break; }
        default: inten = __uint_as_float(load_u32(q, swap)); break;                        // 7 = FLOAT32
      }
    }
    o[4] = inten; o[5] = 0.f; o[6] = 0.f; o[7] = 0.f;
  }
}

// ---- binary PCD header (the files KeyFrame::save writes with pcl::io::savePCDFileBinary, keyframe.cpp:57)
struct PcdHeader {
  std::vector<std::string> fields;
  std::vector<int> size, count;
  std::vector<char> type;
  size_t points = 0, width = 0, height = 1, data_offset = 0;
  std::string data;  // "binary", "ascii", "binary_compressed"
};

inline bool pcd_parse_header(FILE* f, PcdHeader& H) {
  char line[4096];
  long pos = 0;
  bool have_points = false;
  while (fgets(line, sizeof(line), f)) {
    pos = ftell(f);
    std::string s(line);
    while (!s.empty() && (s.back() == '\n' || s.back() == '\r' || s.back() == ' ')) s.pop_back();
    if (s.empty() || s[0] == '#') continue;
    std::vector<std::string> tok;
    size_t a = 0;
    while (a < s.size()) {
      while (a < s.size() && (s[a] == ' ' || s[a] == '\t')) a++;
      size_t b = a;
      while (b < s.size() && s[b] != ' ' && s[b] != '\t') b++;
      if (b > a) tok.push_back(s.substr(a, b - a));
      a = b;
    }
    if (tok.empty()) continue;
    const std::string& k = tok[0];
    if (k == "FIELDS" || k == "COLUMNS") H.fields.assign(tok.begin() + 1, tok.end());
    else if (k == "SIZE") { H.size.clear(); for (size_t i = 1; i < tok.size(); i++) H.size.push_back(atoi(tok[i].c_str())); }
    else if (k == "TYPE") { H.type.clear(); for (size_t i = 1; i < tok.size(); i++) H.type.push_back(tok[i][0]); }
    else if (k == "COUNT") { H.count.clear(); for (size_t i = 1; i < tok.size(); i++) H.count.push_back(atoi(tok[i].c_str())); }
    else if (k == "WIDTH" && tok.size() > 1) H.width = strtoull(tok[1].c_str(), nullptr, 10);
    else if (k == "HEIGHT" && tok.size() > 1) H.height = strtoull(tok[1].c_str(), nullptr, 10);
    else if (k == "POINTS" && tok.size() > 1) { H.points = strtoull(tok[1].c_str(), nullptr, 10); have_points = true; }
    else if (k == "DATA" && tok.size() > 1) { H.data = tok[1]; H.data_offset = (size_t)pos; break; }
  }
  if (!have_points) H.points = H.width * H.height;
  if (H.count.empty()) H.count.assign(H.fields.size(), 1);
  return !H.data.empty() && !H.fields.empty() && H.size.size() == H.fields.size() && H.type.size() == H.fields.size() && H.count.size() == H.fields.size();
}

}  // namespace b2r
