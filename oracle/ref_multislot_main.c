/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * A main() for the UNMODIFIED reference host + CUDA objects (oracle/Makefile.ref, target `multislot`) that drives the reference's
 * multi-slot / adjoint-Jacobian GPU path the way its container front ends do.  The stock command-line program allocates the
 * result volume for `srcnum` sources when the mesh is loaded (src/mmc_mesh.c:654,688) and only later, in mcx_prep
 * (src/mmc_utils.c:3760-3797), appends the detectors as extra source slots; mmc_run_simulation then merges
 * `extrasrclen` slot blocks into that buffer (src/mmc_cu_host.cu:879-881,1003-1024) and the program dies of heap corruption.
 * mmclab/pmmc fill `srcdata` before the mesh is set up, so their buffer has the right size (src/mmc_mesh.c:2389-2394).  This file
 * does the same three calls as src/mmc.c:68-110 and re-sizes the volume in between -- nothing else; kernels, host code,
 * normalisation and mesh_savejacob are the reference's own.
 */
#include <stdlib.h>
#include "mmc_host.h"
#include "mmc_cu_host.h"

int main(int argc, char** argv) {
    mcconfig cfg;
    tetmesh mesh;
    raytracer tracer;

    mmc_init_from_cmd(&cfg, &mesh, &tracer, argc, argv);
    mmc_prep(&cfg, &mesh, &tracer);

    if (cfg.extrasrclen > cfg.srcnum) {
        size_t datalen = (cfg.method == rtBLBadouelGrid) ? (size_t)cfg.crop0.z : (size_t)(cfg.basisorder ? mesh.nn : mesh.ne);

        free(mesh.weight);
        mesh.weight = (double*)calloc(sizeof(double) * datalen * cfg.extrasrclen, cfg.maxgate);
    }

    mmc_run_cu(&cfg, &mesh, &tracer);
    mmc_cleanup(&cfg, &mesh, &tracer);
    return 0;
}
