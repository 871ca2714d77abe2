// fem-shell-coupled -- the coupled program's loop (src/fem-shell/preCICE/fem-shell_precice.cpp,
// main: fsp.cpp:18-419) with a built-in synthetic fluid in place of the preCICE link.
//
//   fem-shell-coupled -nu <v> -e <v> -t <v> -mesh <file.xda> -dt <v> [-axis x|y|z] [-config <xml>]
//                     [-steps n] [-subiters k] [-out <name>] [-d 1] [-pc_type ..] [-ksp_rtol ..] [-ksp_max_it ..]
//
// preCICE itself (library, XML configuration, TCP m2n) is out of scope and not installable here; what is
// kept is the DATA CONTRACT of the structure participant: interface nodes = boundary ids 2/20/21
// (fsp.cpp:55-71), forces in on those nodes (2-D or 3-D, dead axis mapping fsp.cpp:1400-1432), full
// re-solve per coupling iteration, displacement INCREMENT since the last converged step out
// (fsp.cpp:286-317), preSols updated only when the time step converges (fsp.cpp:331-374).
// The fluid stand-in is the reference's own dummy load (fluid_solver.cpp:187-214): 1 + sin(t/25.01) on the
// first coupling component of the first 21 interface nodes; "-subiters k" replays k implicit-coupling
// iterations per time step before the step is committed.
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/femshell_app.hpp"

int main(int argc, char **argv)
{
    double nu = 0.3, em = 1e6, thickness = 1.0, dt = 0.01, rtol = 1e-12;
    long long max_it = 5000;
    std::string mesh_file, out, config, pc = "jacobi";
    char axis = '0';
    int steps = 10, subiters = 1, device = 0;
    bool debug = false, has[5] = {false, false, false, false, false};
    if (argc < 7) {  // fsp.cpp:430-443
        std::cerr << "Error, must choose valid parameters.\nUsage: " << argv[0] << " -nu -e -t -mesh -config -dt [-axis] [-out] [-d]\n";
        return -1;
    }
    for (int i = 1; i + 1 < argc; i++) {
        std::string k = argv[i];
        const char *v = argv[i + 1];
        if (k == "-nu") { nu = atof(v); has[0] = true; }
        else if (k == "-e") { em = atof(v); has[1] = true; }
        else if (k == "-t") { thickness = atof(v); has[2] = true; }
        else if (k == "-mesh") { mesh_file = v; has[3] = true; }
        else if (k == "-dt") { dt = atof(v); has[4] = true; }
        else if (k == "-config") config = v;
        else if (k == "-axis") axis = v[0];
        else if (k == "-out") out = v;
        else if (k == "-d") debug = atoi(v) == 1;
        else if (k == "-steps") steps = atoi(v);
        else if (k == "-subiters") subiters = atoi(v);
        else if (k == "-pc_type") pc = v;
        else if (k == "-ksp_rtol") rtol = atof(v);
        else if (k == "-ksp_max_it") max_it = atoll(v);
        else if (k == "-device") device = atoi(v);
        else continue;
        i++;
    }
    for (bool h : has)
        if (!h) { std::cerr << "ERROR: -nu -e -t -mesh -dt are required\n"; return -1; }
    const int dims = (axis == 'x' || axis == 'y' || axis == 'z') ? 2 : 3;  // 2-D coupling needs a dead axis (fsp.cpp:88-98)

    try {
        fs::app::Mesh mesh;
        mesh.read(mesh_file);
        fs::app::EquationSystems es(mesh, nu, em, thickness, device);
        fs_solve_opts &o = es.solver_options();
        o.rtol = rtol;
        o.max_its = max_it;
        o.pc = pc == "pbjacobi" ? FS_PC_BJACOBI6 : (pc == "none" ? FS_PC_NONE : (pc == "mg" ? FS_PC_MLRBM : FS_PC_JACOBI));
        if (o.pc == FS_PC_MLRBM) o.norm_type = FS_NORM_UNPRECONDITIONED;
        es.init();
        int64_t n_if = 0;
        es.check(fs_interface_nodes(es.context(), &n_if, nullptr));
        std::vector<int32_t> if_nodes(n_if);
        es.check(fs_interface_nodes(es.context(), &n_if, if_nodes.data()));
        std::cout << "coupling dimensions = " << dims << ", dead axis = " << axis << ", coupling interface nodes = " << n_if
                  << ", dt = " << dt << (config.empty() ? "" : ", config = " + config) << std::endl;
        if (n_if == 0) { std::cerr << "no interface nodes (boundary ids 2, 20, 21) in the mesh\n"; return -1; }

        std::vector<double> forces(dims * n_if, 0.0), displ(dims * n_if, 0.0);
        fs_solve_info info;
        for (int t = 0; t < steps; t++) {
            for (int it = 0; it < subiters; it++) {
                for (int64_t i = 0; i < n_if && i < 21; i++) forces[i * dims] = 1.0 + std::sin(t / 25.01);  // fluid_solver.cpp:189-194
                int rc = fs_step(es.context(), dims, dims == 2 ? axis : 'z', forces.data(), &o, displ.data(), &info);
                if (rc != FS_OK && rc != FS_ERR_NOT_CONVERGED) es.check(rc);
                if (debug) {
                    std::cout << "Displacements sent to preCICE:" << std::endl;
                    for (int64_t i = 0; i < n_if; i++) {
                        std::cout << "[" << displ[i * dims];
                        for (int d = 1; d < dims; d++) std::cout << ", " << displ[i * dims + d];
                        std::cout << "]" << std::endl;
                    }
                }
                if (it + 1 < subiters) std::cout << "Iterate" << std::endl;
            }
            es.check(fs_commit_step(es.context(), dims, dims == 2 ? axis : 'z'));
            double mx = 0.0;
            for (double v : displ) mx = std::max(mx, std::fabs(v));
            std::cout << "Advancing in time, finished timestep: " << t << " (" << info.iterations << " CG iterations, "
                      << info.solve_ms << " ms, max |increment| " << mx << ")" << std::endl;
            if (!out.empty()) {
                std::vector<double> sols;
                es.build_solution_vector(sols);
                fs::app::write_vtk(out + "_" + std::to_string(t + 1) + ".vtk", mesh, sols);
            }
        }
        std::cout << "Exiting Structure Solver" << std::endl << "All done :)\n";
        return 0;
    } catch (const fs::app::Error &e) {
        std::cerr << "fem-shell-coupled: " << e.what() << std::endl;
        return -1;
    }
}
