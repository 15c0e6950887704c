// CPU-only check of the file / wire formats in host/pcd_io.hpp (no GPU, no libb200reg): area-list CSV, TUM trajectory,
// PointCloud2 body.  Prints "formats ok" and exits 0, or says what differs.
#include <cstdio>
#include <cmath>
#include "pcd_io.hpp"
using namespace b200host;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)
int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    // area list: write, read back, text format = std::to_string ("%f")
    AreaList areas = {{"surf/tile_0_0.pcd", -50.0, -40.0, -1.5, 50.0, 40.0, 8.25}, {"surf/tile_1_0.pcd", 50.0, -40.0, -1.5, 150.0, 40.0, 8.25}};
    write_arealist(dir + "/arealist.csv", areas);
    {
        std::ifstream f(dir + "/arealist.csv");
        std::string l;
        std::getline(f, l);
        CHECK(l == "surf/tile_0_0.pcd,-50.000000,-40.000000,-1.500000,50.000000,40.000000,8.250000");
    }
    AreaList back = read_arealist(dir + "/arealist.csv");
    CHECK(back.size() == 2 && back[1].path == "surf/tile_1_0.pcd" && back[1].x_max == 150.0 && back[0].z_max == 8.25);
    CHECK(is_in_area(49.0, 0.0, back[0], 0.0) && !is_in_area(51.0, 0.0, back[0], 0.0) && is_in_area(51.0, 0.0, back[0], 5.0));
    std::vector<std::string> near = areas_near(48.0, 0.0, back, "/maps/", 5.0);   // within 5 m of the seam: both tiles
    CHECK(near.size() == 2 && near[0] == "/maps/surf/tile_0_0.pcd");
    CHECK(areas_near(0.0, 0.0, back, "/maps/", 5.0).size() == 1);
    // create_pcd: two tiles on disk, concatenated in list order
    PointXYZI a[2] = {{1, 2, 3, 4}, {5, 6, 7, 8}}, b[1] = {{9, 10, 11, 12}};
    std::system(("mkdir -p " + dir + "/surf").c_str());
    save_pcd_binary(dir + "/surf/tile_0_0.pcd", a, 2);
    save_pcd_binary(dir + "/surf/tile_1_0.pcd", b, 1);
    std::vector<PointXYZI> cat = create_pcd(48.0, 0.0, back, dir + "/", 5.0);
    CHECK(cat.size() == 3 && cat[0].x == 1 && cat[2].intensity == 12);
    // TUM trajectory
    save_trajectory_tum(dir + "/traj.txt", {{1634567890.123456, 1.5, -2.25, 0.125, 0.0, 0.0, 0.38268343236508978, 0.92387953251128674}});
    {
        std::ifstream f(dir + "/traj.txt");
        std::string l0, l1;
        std::getline(f, l0);
        std::getline(f, l1);
        CHECK(l0 == "#timestamp x y z q_x q_y q_z q_w");
        CHECK(l1 == "1634567890.123456 1.500000000000000 -2.250000000000000 0.125000000000000 0.000000000000000 0.000000000000000 0.382683432365090 0.923879532511287");
    }
    // PointCloud2 body of a PointXYZINormal cloud
    const float xyz[2][3] = {{1.f, 2.f, 3.f}, {4.f, 5.f, 6.f}};
    const float inten[2] = {0.5f, 0.75f};
    PointCloud2Blob pc = pack_pointcloud2_xyzinormal(&xyz[0][0], 2, 12, inten);
    CHECK(pc.height == 1 && pc.width == 2 && pc.point_step == 48 && pc.row_step == 96 && pc.data.size() == 96 && pc.is_dense && !pc.is_bigendian);
    CHECK(pc.fields.size() == 8 && std::string(pc.fields[3].name) == "intensity" && pc.fields[3].offset == 32 && pc.fields[7].offset == 36);
    float v;
    std::memcpy(&v, &pc.data[48 + 8], 4);
    CHECK(v == 6.f);
    std::memcpy(&v, &pc.data[48 + 32], 4);
    CHECK(v == 0.75f);
    std::printf("formats ok\n");
    return 0;
}
