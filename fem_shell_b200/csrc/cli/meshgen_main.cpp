// meshGen -- command-line twin of src/meshgen/main_all.cpp on top of fs_meshgen / fs_write_xda:
//   meshGen type nx ny min_x min_y max_x max_y bcids factor loading ul_lr dead-axis filename
// writes <filename>.xda and, when loading > 0, <filename>_f (n header, factor line, n-1 unit-load rows,
// exactly as main_all.cpp:343-387 does).
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/femshell_b200.h"

int main(int argc, char **argv)
{
    if (argc != 14) {  // main_all.cpp:18-35
        std::cout << "usage: " << argv[0] << " type nx ny min_x min_y max_x max_y bcids factor loading ul_lr dead-axis filename\n";
        return -1;
    }
    const char kind = argv[1][0];
    const int nx = atoi(argv[2]), ny = atoi(argv[3]);
    const double min_x = atof(argv[4]), min_y = atof(argv[5]), max_x = atof(argv[6]), max_y = atof(argv[7]);
    int bcids[4] = {-1, -1, -1, -1};
    {
        std::string s = argv[8];
        size_t pos = 0;
        for (int k = 0; k < 4; k++) {
            size_t c = s.find(',', pos);
            bcids[k] = atoi(s.substr(pos, c == std::string::npos ? std::string::npos : c - pos).c_str());
            if (c == std::string::npos) break;
            pos = c + 1;
        }
    }
    const double factor = atof(argv[9]);
    const int loading = atoi(argv[10]);
    const int ul_lr = atoi(argv[11]) == 1;
    const char dead = argv[12][0];
    const std::string name = argv[13];

    int64_t nn = 0, ne = 0, nb = 0;
    int rc = fs_meshgen(kind, nx, ny, min_x, min_y, max_x, max_y, bcids, factor, loading, ul_lr, dead, &nn, &ne, &nb,
                        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc) {
        std::cout << "Invalid arguments (element type Q|q|T|t, positive nx ny, dead axis x|y|z)\n";
        return -1;
    }
    const int nen = (kind == 't' || kind == 'T') ? 3 : 4;
    std::vector<double> xyz(3 * nn), F(6 * nn);
    std::vector<int32_t> etype(ne), enodes(nen * ne), bc(3 * nb + 1);
    std::vector<int64_t> eptr(ne + 1);
    rc = fs_meshgen(kind, nx, ny, min_x, min_y, max_x, max_y, bcids, factor, loading, ul_lr, dead, &nn, &ne, &nb, xyz.data(),
                    etype.data(), eptr.data(), enodes.data(), bc.data(), F.data());
    if (rc) return -1;
    if (fs_write_xda((name + ".xda").c_str(), nn, xyz.data(), ne, etype.data(), eptr.data(), enodes.data(), nb, bc.data())) return -1;
    if (loading <= 0) return 0;
    FILE *f = fopen((name + "_f").c_str(), "w");
    if (!f) return -1;
    const int comp = dead == 'x' ? 0 : (dead == 'y' ? 1 : 2);
    fprintf(f, "%lld\n", (long long)nn);
    if (loading == 1) {
        fprintf(f, "%g\n", factor);
        for (int64_t i = 0; i < nn - 1; i++) {
            int v[6] = {0, 0, 0, 0, 0, 0};
            if (i == nn / 2) v[comp] = 1;
            fprintf(f, "%d %d %d %d %d %d\n", v[0], v[1], v[2], v[3], v[4], v[5]);
        }
    } else if (loading == 2) {
        fprintf(f, "%g\n", factor * ((max_x - min_x) / (double)nx) * ((max_y - min_y) / (double)ny));
        int v[6] = {0, 0, 0, 0, 0, 0};
        v[comp] = 1;
        for (int64_t i = 0; i < nn - 1; i++) fprintf(f, "%d %d %d %d %d %d\n", v[0], v[1], v[2], v[3], v[4], v[5]);
    }
    fclose(f);
    return 0;
}
