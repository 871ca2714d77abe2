// fem-shell (stand-alone) -- command-line twin of the reference program src/fem-shell/fem-shell.cpp
// (main: fs.cpp:14-185, flags: fs.cpp:194-267) on top of the B200 library.
//
//   fem-shell -nu <v> -e <v> -t <v> -mesh <file.xda> [-out <name>] [-d 1]
//             [-ksp_type cg] [-pc_type jacobi|pbjacobi|none|mg] [-ksp_rtol r] [-ksp_max_it n]
//             [-ksp_norm_type preconditioned|unpreconditioned] [-dof_order libmesh|node] [-device k]
//
// Flags after the reference's own six are the PETSc options the reference passes through to KSP
// (doc/implementation.tex:68-72).  Differences from the reference, stated once: the Krylov method is
// always CG (north_star), and -out writes a legacy .vtk instead of ExodusII.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/femshell_app.hpp"

namespace {

struct Args {
    bool debug = false, has_nu = false, has_e = false, has_t = false, has_mesh = false, has_out = false;
    double nu = 0.3, em = 1.0e6, thickness = 1.0;
    std::string mesh, out;
    std::string pc = "jacobi", norm = "preconditioned", dof = "libmesh";
    double rtol = 1e-12;
    long long max_it = 5000;
    int device = 0;
};

// GetPot-style: "-flag value" pairs in any order (fs.cpp:212-255)
bool parse(int argc, char **argv, Args &a)
{
    for (int i = 1; i < argc; i++) {
        std::string k = argv[i];
        const char *v = (i + 1 < argc) ? argv[i + 1] : nullptr;
        auto take = [&]() { i++; return v; };
        if (k == "-d" && v) a.debug = atoi(take()) == 1;
        else if (k == "-nu" && v) { a.nu = atof(take()); a.has_nu = true; }
        else if (k == "-e" && v) { a.em = atof(take()); a.has_e = true; }
        else if (k == "-t" && v) { a.thickness = atof(take()); a.has_t = true; }
        else if (k == "-mesh" && v) { a.mesh = take(); a.has_mesh = true; }
        else if (k == "-out" && v) { a.out = take(); a.has_out = true; }
        else if (k == "-pc_type" && v) a.pc = take();
        else if (k == "-ksp_rtol" && v) a.rtol = atof(take());
        else if (k == "-ksp_max_it" && v) a.max_it = atoll(take());
        else if (k == "-ksp_norm_type" && v) a.norm = take();
        else if (k == "-ksp_type" && v) { if (std::string(take()) != "cg") std::cerr << "note: only -ksp_type cg is implemented; using cg\n"; }
        else if (k == "-dof_order" && v) a.dof = take();
        else if (k == "-device" && v) a.device = atoi(take());
    }
    bool failed = false;
    if (!a.has_nu) { std::cerr << "ERROR: Poisson's ratio nu not specified!\n"; failed = true; }
    if (!a.has_e) { std::cerr << "ERROR: Elastic modulus E not specified!\n"; failed = true; }
    if (!a.has_t) { std::cerr << "ERROR: Mesh thickness t not specified!\n"; failed = true; }
    if (!a.has_mesh) { std::cerr << "ERROR: Mesh file not specified!\n"; failed = true; }
    return !failed;
}

}  // namespace

int main(int argc, char **argv)
{
    Args a;
    if (argc < 5) {  // fs.cpp:196-208
        std::cerr << "Error, must choose valid parameters.\n"
                  << "Usage: " << argv[0] << " -nu -e -t -mesh [-out] [-d]\n"
                  << "-nu:\t Possion's ratio (required)\n"
                  << "-e:\t Elastic/Young's modulus E (required)\n"
                  << "-t:\t Thickness (required)\n"
                  << "-mesh:\t Input mesh file (*.xda, required)\n"
                  << "-out:\t Output file name (without extension, optional)\n"
                  << "-d:\t Additional (debug) messages (1=on, 0=off (default))\n";
        std::cout << "Read command-line arguments.......FAILED" << std::endl;
        return -1;
    }
    if (!parse(argc, argv, a)) {
        std::cout << "Read command-line arguments.......FAILED" << std::endl;
        return -1;
    }
    std::cout << "Run program with parameters:"
              << " debug messages = " << (a.debug ? "true" : "false") << ", nu = " << a.nu << ", E = " << a.em
              << ", t = " << a.thickness << ", mesh file = " << a.mesh;
    if (a.has_out) std::cout << ", out-file = " << a.out;
    std::cout << std::endl;
    std::cout << "Read command-line arguments.......OK" << std::endl;

    try {
        fs::app::Mesh mesh;
        mesh.read(a.mesh);
        std::cout << " Mesh Information:\n  n_nodes()=" << mesh.n_nodes() << "\n  n_elem()=" << mesh.n_elem() << std::endl;
        mesh.read_forces();

        fs::app::EquationSystems es(mesh, a.nu, a.em, a.thickness, a.device);
        es.set_dof_order(a.dof == "node" ? FS_DOF_NODE_ID : FS_DOF_FIRST_ENCOUNTER);
        fs_solve_opts &o = es.solver_options();
        o.rtol = a.rtol;
        o.max_its = a.max_it;
        o.pc = a.pc == "pbjacobi" ? FS_PC_BJACOBI6 : (a.pc == "none" ? FS_PC_NONE : (a.pc == "mg" ? FS_PC_MLRBM : FS_PC_JACOBI));
        // the multilevel cycle is not a fixed diagonal scaling: only the true residual norm is defined for it
        o.norm_type = (a.norm == "unpreconditioned" || o.pc == FS_PC_MLRBM) ? FS_NORM_UNPRECONDITIONED : FS_NORM_PRECONDITIONED;
        es.init();
        int64_t n_dofnodes = 0, n_blocks = 0;
        fs_get_sizes(es.context(), &n_dofnodes, &n_blocks, nullptr, nullptr, nullptr);
        std::cout << " EquationSystems\n  System \"Elasticity\"\n   n_dofs()=" << 6 * n_dofnodes << "\n   n_nonzeros=" << 36 * n_blocks << std::endl;

        int status = es.solve();
        const fs_solve_info &info = es.last_solve();
        std::cout << "Linear solver: cg/" << a.pc << ", " << info.iterations << " iterations, relative residual "
                  << info.rel_residual << (status == FS_OK ? "" : " (NOT CONVERGED)") << ", assembly " << es.assemble_ms()
                  << " ms, solve " << info.solve_ms << " ms" << std::endl;
        std::vector<double> sols;
        es.build_solution_vector(sols);

        if (a.debug) {  // fs.cpp:143-150
            std::vector<int64_t> rowptr(6 * n_dofnodes + 1);
            std::vector<int32_t> col(36 * n_blocks);
            std::vector<double> val(36 * n_blocks), rhs(6 * n_dofnodes);
            es.check(fs_export_csr(es.context(), rowptr.data(), col.data(), val.data()));
            es.check(fs_export_rhs(es.context(), rhs.data()));
            std::cout << "System matrix:" << std::endl;
            for (int64_t r = 0; r < 6 * n_dofnodes; r++) {
                std::cout << "row " << r << ":";
                for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) std::cout << " (" << col[k] << ", " << val[k] << ") ";
                std::cout << "\n";
            }
            std::cout << std::endl << "RHS:" << std::endl;
            for (double v : rhs) std::cout << v << "\n";
            std::cout << std::endl;
        }

        // fs.cpp:156-176
        std::cout << "Solution: u_vec = [";
        for (int64_t id = 0; id < mesh.n_nodes(); id++) {
            std::cout << "u= " << sols[6 * id] << ", v= " << sols[6 * id + 1] << ", w= " << sols[6 * id + 2];
            std::cout << ", tx= " << sols[6 * id + 3] << ", ty= " << sols[6 * id + 4] << ", tz= " << sols[6 * id + 5] << "]" << std::endl;
        }
        std::cout << "]" << std::endl << std::endl;

        if (a.has_out) {  // nodal displacements + element stress resultants (doc/shellelements.tex:1394-1403)
            std::vector<double> res;
            es.build_resultants(res);
            if (!fs::app::write_vtk(a.out + ".vtk", mesh, sols, &res)) std::cerr << "could not write " << a.out << ".vtk\n";
        }
        std::cout << "All done :)\n";
        return status == FS_OK ? 0 : 1;
    } catch (const fs::app::Error &e) {
        std::cerr << "fem-shell: " << e.what() << std::endl;
        return -1;
    }
}
