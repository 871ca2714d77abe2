// femshell_app.hpp -- header-only C++ convenience layer over the C ABI (include/femshell_b200.h) that
// mirrors the call shape of the reference's libMesh program (src/fem-shell/fem-shell.cpp):
//
//   reference (fs.cpp)                                   here
//   Mesh mesh(comm, 2); mesh.read(file)        :35-37    fs::app::Mesh mesh; mesh.read(file)
//   forces <- "<base>_f"                       :44-67    mesh.read_forces()
//   EquationSystems es(mesh); add_system       :70-83    fs::app::EquationSystems es(mesh, nu, E, t)
//   system.attach_assemble_function(f)         :85       es.attach_assemble_function(f)   (optional hook)
//   es.init()                                  :125      es.init()
//   es.solve()                                 :138      es.solve()          (assemble + Krylov solve)
//   es.build_solution_vector(sols)             :141      es.build_solution_vector(sols)
//
// It adds no numerics: every method is one or two C-ABI calls.
#pragma once
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "femshell_b200.h"

namespace fs {
namespace app {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

struct Mesh {
    std::vector<double> xyz;
    std::vector<int32_t> etype, enodes, bc;
    std::vector<int64_t> eptr;
    std::vector<double> forces;  // 6 per node, already scaled by the global factor
    std::string file;

    int64_t n_nodes() const { return (int64_t)xyz.size() / 3; }
    int64_t n_elem() const { return (int64_t)etype.size(); }

    void read(const std::string &path)
    {
        file = path;
        int64_t nn = 0, ne = 0, nen = 0, nb = 0;
        int rc = fs_read_mesh(path.c_str(), &nn, &ne, &nen, &nb, nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc) throw Error(rc, "cannot read mesh file " + path);
        xyz.resize(3 * nn); etype.resize(ne); eptr.resize(ne + 1); enodes.resize(nen); bc.resize(3 * nb);
        rc = fs_read_mesh(path.c_str(), &nn, &ne, &nen, &nb, xyz.data(), etype.data(), eptr.data(), enodes.data(), bc.data());
        if (rc) throw Error(rc, "cannot read mesh file " + path);
        forces.assign(6 * nn, 0.0);
    }

    // CONVENTION (fs.cpp:41-50): force file = mesh file name without extension + "_f"; a missing file means no loads
    bool read_forces()
    {
        std::string f = file;
        if (f.find(".xda") != std::string::npos || f.find(".xdr") != std::string::npos || f.find(".msh") != std::string::npos)
            f.resize(f.size() - 4);
        f += "_f";
        forces.assign(6 * n_nodes(), 0.0);
        return fs_read_forces(f.c_str(), n_nodes(), forces.data()) == FS_OK;
    }
};

class EquationSystems {
public:
    using AssembleHook = std::function<void(EquationSystems &, const std::string &)>;

    EquationSystems(Mesh &mesh, double nu, double E, double t, int device = 0) : mesh_(mesh)
    {
        int rc = fs_create(&ctx_, device);
        if (rc) throw Error(rc, "fs_create failed: no usable CUDA device (there is no CPU fallback)");
        check(fs_set_material(ctx_, nu, E, t));
        opts_.rtol = 1e-12;  // libMesh defaults the reference leaves untouched (fs.cpp:130-133)
        opts_.max_its = 5000;
        opts_.pc = FS_PC_JACOBI;
        opts_.norm_type = FS_NORM_PRECONDITIONED;
        opts_.warm_start = 1;
        opts_.check_every = 0;
    }
    ~EquationSystems() { if (ctx_) fs_destroy(ctx_); }
    EquationSystems(const EquationSystems &) = delete;
    EquationSystems &operator=(const EquationSystems &) = delete;

    fs_context *context() { return ctx_; }
    fs_solve_opts &solver_options() { return opts_; }
    const fs_solve_info &last_solve() const { return info_; }
    void set_dof_order(int mode) { check(fs_set_dof_order(ctx_, mode)); }

    // optional user hook run right before the built-in assembly, with the signature of the reference's callback
    void attach_assemble_function(AssembleHook h) { hook_ = std::move(h); }

    void init()
    {
        check(fs_set_mesh(ctx_, mesh_.n_nodes(), mesh_.xyz.data(), mesh_.n_elem(), mesh_.etype.data(), mesh_.eptr.data(),
                          mesh_.enodes.data(), (int64_t)mesh_.bc.size() / 3, mesh_.bc.data()));
        check(fs_set_nodal_loads(ctx_, mesh_.forces.data()));
    }

    // LinearImplicitSystem::solve: zero K,b -> assemble -> Krylov solve.  Returns the solver status
    // (FS_OK or FS_ERR_NOT_CONVERGED); other errors throw.
    int solve(bool reassemble = true)
    {
        if (hook_) hook_(*this, "Elasticity");
        if (reassemble || !assembled_) { check(fs_assemble(ctx_, &assemble_ms_)); assembled_ = true; }
        int rc = fs_solve(ctx_, &opts_, &info_);
        if (rc != FS_OK && rc != FS_ERR_NOT_CONVERGED) check(rc);
        return rc;
    }

    void build_solution_vector(std::vector<double> &sols)
    {
        sols.resize(6 * mesh_.n_nodes());
        check(fs_get_solution(ctx_, sols.data()));
    }

    // membrane stresses and bending moments at the element centroids (fs_recover_resultants): 6 per element
    void build_resultants(std::vector<double> &res)
    {
        res.resize(6 * mesh_.n_elem());
        check(fs_recover_resultants(ctx_, res.data()));
    }

    float assemble_ms() const { return assemble_ms_; }

    void check(int rc)
    {
        if (rc != FS_OK) throw Error(rc, fs_last_error(ctx_));
    }

private:
    Mesh &mesh_;
    fs_context *ctx_ = nullptr;
    fs_solve_opts opts_;
    fs_solve_info info_ = {};
    AssembleHook hook_;
    bool assembled_ = false;
    float assemble_ms_ = 0.f;
};

// legacy-VTK (ASCII, unstructured grid) writer of the displaced mesh with the six nodal fields; stands in for the
// reference's ExodusII / VTK output (fs.cpp:1240-1251, fsp.cpp:1526-1561), whose libraries are not available here
// `resultants` (optional, 6 per element from fs_recover_resultants) is written as cell data
inline bool write_vtk(const std::string &path, const Mesh &m, const std::vector<double> &sols,
                      const std::vector<double> *resultants = nullptr)
{
    FILE *f = fopen(path.c_str(), "w");
    if (!f) return false;
    const int64_t nn = m.n_nodes(), ne = m.n_elem();
    fprintf(f, "# vtk DataFile Version 3.0\nfem-shell displaced mesh\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS %lld double\n", (long long)nn);
    for (int64_t i = 0; i < nn; i++)  // fs.cpp:172-174: nodes moved by (u,v,w)
        fprintf(f, "%.17g %.17g %.17g\n", m.xyz[3 * i] + sols[6 * i], m.xyz[3 * i + 1] + sols[6 * i + 1], m.xyz[3 * i + 2] + sols[6 * i + 2]);
    fprintf(f, "CELLS %lld %lld\n", (long long)ne, (long long)(ne + m.eptr[ne]));
    for (int64_t e = 0; e < ne; e++) {
        fprintf(f, "%d", (int)(m.eptr[e + 1] - m.eptr[e]));
        for (int64_t k = m.eptr[e]; k < m.eptr[e + 1]; k++) fprintf(f, " %d", m.enodes[k]);
        fputc('\n', f);
    }
    fprintf(f, "CELL_TYPES %lld\n", (long long)ne);
    for (int64_t e = 0; e < ne; e++) fprintf(f, "%d\n", m.etype[e] == FS_TRI3 ? 5 : 9);
    fprintf(f, "POINT_DATA %lld\n", (long long)nn);
    static const char *names[6] = {"u", "v", "w", "tx", "ty", "tz"};
    for (int v = 0; v < 6; v++) {
        fprintf(f, "SCALARS %s double 1\nLOOKUP_TABLE default\n", names[v]);
        for (int64_t i = 0; i < nn; i++) fprintf(f, "%.17g\n", sols[6 * i + v]);
    }
    if (resultants && (int64_t)resultants->size() == 6 * ne) {
        static const char *rnames[6] = {"sigma_xx", "sigma_yy", "sigma_xy", "M_x", "M_y", "M_xy"};
        fprintf(f, "CELL_DATA %lld\n", (long long)ne);
        for (int v = 0; v < 6; v++) {
            fprintf(f, "SCALARS %s double 1\nLOOKUP_TABLE default\n", rnames[v]);
            for (int64_t e = 0; e < ne; e++) fprintf(f, "%.17g\n", (*resultants)[6 * e + v]);
        }
    }
    fclose(f);
    return true;
}

}  // namespace app
}  // namespace fs
