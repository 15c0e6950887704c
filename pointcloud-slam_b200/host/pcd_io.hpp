// b200reg host adaptor — the on-disk formats of the map pipeline, dependency-free.
//
//   frames/<int>.pcd      pcl::PointXYZI keyframes written by pcl::io::savePCDFileBinary (also ascii), enumerated by the
//                         numeric value of the basename (tool/occupancy_mapping/src/mapping_server.cc:500-538)
//   poses_{ori,opt}.txt   one pose per line: x y z qw qx qy qz (mapping_server.cc:466-497)
//   jueying.pcd           the merged map, binary PointXYZI (jueying_lio/src/laser_mapping.cc:887-897 writes the LIO map the same way)
// Only the fields x, y, z and intensity are read; any other field (normals, curvature, ring, time) is skipped by its
// declared SIZE * COUNT, so PointXYZINormal and driver-specific clouds load too.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200host {

struct PointXYZI { float x, y, z, intensity; };

inline std::vector<PointXYZI> load_pcd(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::vector<std::string> fields;
    std::vector<int> size, count;
    std::vector<char> type;
    size_t points = 0, width = 0, height = 1;
    std::string data, line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key;
        ss >> key;
        if (key == "FIELDS" || key == "COLUMNS") { std::string f; while (ss >> f) fields.push_back(f); }
        else if (key == "SIZE") { int v; while (ss >> v) size.push_back(v); }
        else if (key == "TYPE") { char v; while (ss >> v) type.push_back(v); }
        else if (key == "COUNT") { int v; while (ss >> v) count.push_back(v); }
        else if (key == "WIDTH") ss >> width;
        else if (key == "HEIGHT") ss >> height;
        else if (key == "POINTS") ss >> points;
        else if (key == "DATA") { ss >> data; break; }
    }
    if (fields.empty() || size.size() != fields.size() || type.size() != fields.size()) throw std::runtime_error("bad PCD header in " + path);
    if (count.empty()) count.assign(fields.size(), 1);
    if (!points) points = width * height;
    int off[4] = {-1, -1, -1, -1}, fidx[4] = {-1, -1, -1, -1};
    int stride = 0, col = 0;
    for (size_t f = 0; f < fields.size(); ++f) {
        const char* names[4] = {"x", "y", "z", "intensity"};
        for (int k = 0; k < 4; ++k)
            if (fields[f] == names[k]) {
                if (size[f] != 4 || type[f] != 'F') throw std::runtime_error("field " + fields[f] + " is not float32 in " + path);
                off[k] = stride;
                fidx[k] = col;
            }
        stride += size[f] * count[f];
        col += count[f];
    }
    if (off[0] < 0 || off[1] < 0 || off[2] < 0) throw std::runtime_error("PCD without x y z: " + path);
    std::vector<PointXYZI> out(points);
    if (data == "binary") {
        std::vector<char> buf(points * (size_t)stride);
        in.read(buf.data(), (std::streamsize)buf.size());
        if ((size_t)in.gcount() != buf.size()) throw std::runtime_error("truncated PCD " + path);
        for (size_t i = 0; i < points; ++i) {
            const char* p = buf.data() + i * stride;
            memcpy(&out[i].x, p + off[0], 4);
            memcpy(&out[i].y, p + off[1], 4);
            memcpy(&out[i].z, p + off[2], 4);
            if (off[3] >= 0) memcpy(&out[i].intensity, p + off[3], 4); else out[i].intensity = 0.f;
        }
    } else if (data == "ascii") {
        for (size_t i = 0; i < points; ++i) {
            if (!std::getline(in, line)) throw std::runtime_error("truncated PCD " + path);
            std::istringstream ss(line);
            std::vector<double> v;
            double d;
            while (ss >> d) v.push_back(d);
            if ((int)v.size() < col) throw std::runtime_error("short row in " + path);
            out[i] = PointXYZI{(float)v[fidx[0]], (float)v[fidx[1]], (float)v[fidx[2]], fidx[3] >= 0 ? (float)v[fidx[3]] : 0.f};
        }
    } else {
        throw std::runtime_error("unsupported PCD DATA '" + data + "' (binary_compressed is not handled) in " + path);
    }
    return out;
}

/// pcl::io::savePCDFileBinary of a pcl::PointCloud<pcl::PointXYZI>
inline void save_pcd_binary(const std::string& path, const PointXYZI* pts, size_t n) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
               "WIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", n, n);
    if (n && fwrite(pts, sizeof(PointXYZI), n, f) != n) { fclose(f); throw std::runtime_error("short write " + path); }
    fclose(f);
}

/// poses.txt: x y z qw qx qy qz per line
inline std::vector<std::array<double, 7>> load_poses(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::vector<std::array<double, 7>> poses;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::array<double, 7> p;
        int k = 0;
        while (k < 7 && (ss >> p[k])) ++k;
        if (k == 0) continue;
        if (k != 7) throw std::runtime_error("pose line without 7 numbers in " + path);
        poses.push_back(p);
    }
    return poses;
}

/// <dir>/<int>.pcd in ascending numeric order of the basename
inline std::vector<std::string> list_frames(const std::string& dir) {
    std::vector<std::pair<long, std::string>> v;
    DIR* d = opendir(dir.c_str());
    if (!d) throw std::runtime_error("cannot open directory " + dir);
    while (dirent* e = readdir(d)) {
        const std::string name = e->d_name;
        if (name.size() < 5 || name.substr(name.size() - 4) != ".pcd") continue;
        char* end = nullptr;
        const long id = strtol(name.c_str(), &end, 10);
        if (end == name.c_str()) continue;
        v.push_back({id, dir + "/" + name});
    }
    closedir(d);
    std::sort(v.begin(), v.end());
    std::vector<std::string> out;
    for (auto& p : v) out.push_back(p.second);
    return out;
}

// ------------------------------------------------------------------ tiled prior maps: the area list
// jueying_slam/include/dynamic_map.h:14-106: one CSV line per map tile, `path,x_min,y_min,z_min,x_max,y_max,z_max`
// (read_csv splits on ',' only; write_arealist formats with std::to_string = "%f"), is_in_area (:113-116) selects the
// tiles whose xy box, grown by a margin, contains the pose; create_pcd (:127-156) concatenates them in list order.
struct Area {
    std::string path;
    double x_min = 0, y_min = 0, z_min = 0, x_max = 0, y_max = 0, z_max = 0;
};
using AreaList = std::vector<Area>;

inline AreaList read_arealist(const std::string& path) {
    std::ifstream ifs(path.c_str());
    if (!ifs) throw std::runtime_error("cannot open " + path);
    AreaList ret;
    std::string line;
    while (std::getline(ifs, line)) {
        std::istringstream iss(line);
        std::string col;
        std::vector<std::string> cols;
        while (std::getline(iss, col, ',')) cols.push_back(col);
        if (cols.size() < 7) throw std::runtime_error("arealist line with fewer than 7 columns: " + line);
        Area a;
        a.path = cols[0];
        a.x_min = std::stod(cols[1]); a.y_min = std::stod(cols[2]); a.z_min = std::stod(cols[3]);
        a.x_max = std::stod(cols[4]); a.y_max = std::stod(cols[5]); a.z_max = std::stod(cols[6]);
        ret.push_back(a);
    }
    return ret;
}
inline void write_arealist(const std::string& path, const AreaList& areas) {
    std::ofstream ofs(path.c_str());
    if (!ofs) throw std::runtime_error("cannot write " + path);
    for (const Area& a : areas)
        ofs << a.path << "," << std::to_string(a.x_min) << "," << std::to_string(a.y_min) << "," << std::to_string(a.z_min) << ","
            << std::to_string(a.x_max) << "," << std::to_string(a.y_max) << "," << std::to_string(a.z_max) << std::endl;
}
inline bool is_in_area(double x, double y, const Area& a, double m) {
    return (a.x_min - m) <= x && x <= (a.x_max + m) && (a.y_min - m) <= y && y <= (a.y_max + m);
}
/// the tiles create_pcd(p_x, p_y, areas, global_path, margin) would load, in list order
inline std::vector<std::string> areas_near(double x, double y, const AreaList& areas, const std::string& global_path, double margin) {
    std::vector<std::string> out;
    for (const Area& a : areas)
        if (is_in_area(x, y, a, margin)) out.push_back(global_path + a.path);
    return out;
}
/// create_pcd: the concatenation of those tiles (binary / ascii .pcd), ready for setInputTarget
inline std::vector<PointXYZI> create_pcd(double x, double y, const AreaList& areas, const std::string& global_path, double margin) {
    std::vector<PointXYZI> all;
    for (const std::string& f : areas_near(x, y, areas, global_path, margin)) {
        const std::vector<PointXYZI> part = load_pcd(f);
        all.insert(all.end(), part.begin(), part.end());
    }
    return all;
}

// ------------------------------------------------------------------ trajectory file
// LaserMapping::Savetrajectory (jueying_lio/src/laser_mapping.cc:825-841): TUM text, header line, stamp with 6 decimals,
// the other seven numbers with 15.
struct StampedPose {
    double stamp, x, y, z, qx, qy, qz, qw;
};
inline void save_trajectory_tum(const std::string& path, const std::vector<StampedPose>& poses) {
    std::ofstream ofs(path, std::ios::out);
    if (!ofs.is_open()) throw std::runtime_error("Failed to open traj_file: " + path);
    ofs << "#timestamp x y z q_x q_y q_z q_w" << std::endl;
    for (const StampedPose& p : poses) {
        ofs.setf(std::ios::fixed);
        ofs.precision(6);
        ofs << p.stamp << " ";
        ofs.precision(15);
        ofs << p.x << " " << p.y << " " << p.z << " " << p.qx << " " << p.qy << " " << p.qz << " " << p.qw << std::endl;
    }
}

// ------------------------------------------------------------------ /cloud_registered: sensor_msgs/PointCloud2 payload
// PublishFrameWorld (laser_mapping.cc:747-773) sends pcl::toROSMsg(PointCloud<PointXYZINormal>): height 1, width n,
// point_step 48, little endian, the eight FLOAT32 fields at the offsets of pcl::PointXYZINormal.  The blob below is that
// message body without the ROS header (stamp = lidar_end_time, frame_id "camera_init" are set by the node).
struct PointField {
    const char* name;
    uint32_t offset;
    uint8_t datatype;  // 7 = sensor_msgs::PointField::FLOAT32
    uint32_t count;
};
struct PointCloud2Blob {
    uint32_t height = 1, width = 0, point_step = 48, row_step = 0;
    bool is_bigendian = false, is_dense = true;
    std::vector<PointField> fields;
    std::vector<uint8_t> data;
};
inline const std::vector<PointField>& xyzinormal_fields() {
    static const std::vector<PointField> f = {{"x", 0, 7, 1},         {"y", 4, 7, 1},          {"z", 8, 7, 1},          {"intensity", 32, 7, 1},
                                               {"normal_x", 16, 7, 1}, {"normal_y", 20, 7, 1}, {"normal_z", 24, 7, 1}, {"curvature", 36, 7, 1}};
    return f;
}
/// xyz (+ optional per-point intensity) at `stride` bytes -> the PointCloud2 body of a PointXYZINormal cloud
inline PointCloud2Blob pack_pointcloud2_xyzinormal(const float* xyz, size_t n, size_t stride_bytes, const float* intensity = nullptr) {
    PointCloud2Blob b;
    b.width = (uint32_t)n;
    b.row_step = b.point_step * b.width;
    b.fields = xyzinormal_fields();
    b.data.assign(n * 48, 0);
    for (size_t i = 0; i < n; ++i) {
        const float* p = reinterpret_cast<const float*>(reinterpret_cast<const char*>(xyz) + i * stride_bytes);
        float rec[12] = {p[0], p[1], p[2], 1.0f, 0, 0, 0, 0, intensity ? intensity[i] : 0.0f, 0, 0, 0};  // data[3] = 1 (PCL_ADD_POINT4D)
        std::memcpy(&b.data[i * 48], rec, 48);
        if (!(p[0] == p[0] && p[1] == p[1] && p[2] == p[2])) b.is_dense = false;
    }
    return b;
}

}  // namespace b200host
