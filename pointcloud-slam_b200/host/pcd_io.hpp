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

}  // namespace b200host
