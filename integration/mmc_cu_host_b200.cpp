// Reference-side binding of the mmc_b200 C-ABI: a drop-in replacement for the reference's src/mmc_cu_host.cu.
//
// It exports the one symbol the reference's callers link against,
//     void mmc_run_cu(mcconfig* cfg, tetmesh* mesh, raytracer* tracer)           (src/mmc_cu_host.h:58)
// (called from src/mmc.c:86-90, src/mmclab.cpp:366-370, src/pmmc.cpp:1064-1068 when cfg->compute == cbCUDA), unpacks
// the reference's own structs (mcconfig src/mmc_utils.h:210-345, tetmesh src/mmc_mesh.h:87-122) into the plain-pointer
// structs of include/mmc_b200.h and forwards failures to mcx_error() exactly like CUDA_ASSERT does
// (src/mmc_cu_host.cu:52-53,99-103).  It is compiled AGAINST THE REFERENCE HEADERS, so it lives outside the product
// library; oracle/Makefile.ref target `b200cli` links it with the reference's unmodified host objects into
// oracle/_ref/mmc_b200cli -- the stock `mmc` command line driving the B200 engine (tests/test_cli_dropin.py).
//
// Build (what a maintainer adds to src/Makefile instead of the mmc_cu_host.cu rule):
//     g++ -c -DUSE_CUDA -DMMC_XORSHIFT -DUSE_OS_TIMER -I$(MMC)/src -I$(MMC_B200)/include mmc_cu_host_b200.cpp
//     ... -L$(MMC_B200)/mmc_b200 -lmmc_b200
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "mmc_cu_host.h"        // reference header: mcconfig, tetmesh, raytracer, mcx_error, MMC_FPRINTF
#include "mmc_tictoc.h"
#include "mmc_b200.h"

#ifdef _OPENMP
    #include <omp.h>
#endif

#define B200_ASSERT(rc) b200_assess((rc), __FILE__, __LINE__)

static void b200_assess(int rc, const char* file, int line) {
    if (rc < 0) {
        mcx_error(rc, (char*)mmcb_last_error(), file, line);
    }
}

extern "C" int mcx_list_cu_gpu(mcconfig* cfg, GPUInfo** info) {     // src/mmc_cu_host.cu:108-198
    mmcb_gpuinfo tmp[MAX_DEVICE];
    int count = mmcb_list_gpu(tmp, MAX_DEVICE), activedev = 0;

    if (count <= 0) {
        MMC_FPRINTF(stderr, S_RED "ERROR: No CUDA-capable GPU device found\n" S_RESET);
        return 0;
    }

    *info = (GPUInfo*)calloc(count, sizeof(GPUInfo));

    if (cfg->gpuid && cfg->gpuid > (uint)count) {
        MMC_FPRINTF(stderr, S_RED "ERROR: Specified GPU ID is out of range\n" S_RESET);
        return 0;
    }

    for (int dev = 0; dev < count && dev < MAX_DEVICE; dev++) {
        GPUInfo* g = *info + dev;

        if (cfg->isgpuinfo == 3) {
            activedev++;
        } else if (cfg->deviceid[dev] == '1') {
            cfg->deviceid[dev] = '\0';
            cfg->deviceid[activedev] = dev + 1;
            activedev++;
        }

        strncpy(g->name, tmp[dev].name, MAX_SESSION_LENGTH - 1);
        g->id = tmp[dev].id;
        g->devcount = tmp[dev].devcount;
        g->major = tmp[dev].major;
        g->minor = tmp[dev].minor;
        g->globalmem = tmp[dev].globalmem;
        g->constmem = tmp[dev].constmem;
        g->sharedmem = tmp[dev].sharedmem;
        g->regcount = tmp[dev].regcount;
        g->clock = tmp[dev].clock;
        g->sm = tmp[dev].sm;
        g->core = tmp[dev].core;
        g->autoblock = tmp[dev].autoblock;
        g->autothread = tmp[dev].autothread;
        g->maxgate = cfg->maxgate;
        g->maxmpthread = tmp[dev].maxmpthread;

        if (cfg->isgpuinfo) {
            MMC_FPRINTF(stdout, S_BLUE "=============================   GPU Infomation  ================================\n" S_RESET);
            MMC_FPRINTF(stdout, "Device %d of %d:\t\t%s\n", g->id, g->devcount, g->name);
            MMC_FPRINTF(stdout, "Compute Capability:\t%u.%u\n", g->major, g->minor);
            MMC_FPRINTF(stdout, "Global Memory:\t\t%zu B\nConstant Memory:\t%zu B\nShared Memory:\t\t%zu B\nRegisters:\t\t%u\nClock Speed:\t\t%.2f GHz\n",
                        g->globalmem, g->constmem, g->sharedmem, (unsigned int)g->regcount, g->clock * 1e-6f);
            MMC_FPRINTF(stdout, "Number of MPs:\t\t%u\nNumber of Cores:\t%u\nSMX count:\t\t%u\n", g->sm, g->core, g->sm);
        }
    }

    if (cfg->isgpuinfo == 2 && cfg->parentid == mpStandalone) {
        exit(0);
    }

    if (activedev < MAX_DEVICE) {
        cfg->deviceid[activedev] = '\0';
    }

    return activedev;
}

extern "C" void mmc_run_cu(mcconfig* cfg, tetmesh* mesh, raytracer* tracer) {
    GPUInfo* gpuinfo = NULL;
    unsigned int activedev = 0;
    (void)tracer;               // the engine builds its own tables (96-byte plane records) from node/elem

    if (!(activedev = mcx_list_cu_gpu(cfg, &gpuinfo))) {
        mcx_error(-1, "No GPU device found\n", __FILE__, __LINE__);
    }

    // ---- mesh: plain arrays.  FLOAT3 is 12 bytes under nvcc but 16 bytes in SSE host builds (src/mmc_vector_types.h:69-79)
    std::vector<float> node(3 * (size_t)mesh->nn);
    std::vector<int> elem(4 * (size_t)mesh->ne), facenb;

    for (int i = 0; i < mesh->nn; i++) {
        node[3 * (size_t)i] = mesh->node[i].x;
        node[3 * (size_t)i + 1] = mesh->node[i].y;
        node[3 * (size_t)i + 2] = mesh->node[i].z;
    }

    for (int i = 0; i < mesh->ne; i++)
        for (int j = 0; j < 4; j++) {
            elem[4 * (size_t)i + j] = mesh->elem[(size_t)i * mesh->elemlen + j];
        }

    if (mesh->facenb) {         // tracer_prep numbered the exterior faces -(1..nf) (src/mmc_mesh.c:1466-1474); the ABI takes 0
        facenb.resize(4 * (size_t)mesh->ne);

        for (size_t i = 0; i < facenb.size(); i++) {
            facenb[i] = mesh->facenb[i] > 0 ? mesh->facenb[i] : 0;
        }
    }

    // The callers have already consumed two things the C-ABI takes in their original form:
    //  * labels: mesh_srcdetelem (src/mmc_mesh.c:390-427) moved the -1 labels (wide-field source candidates) into mesh->srcelem and
    //    reset them to 0; mesh_loadmedia / mesh_validate turned the -2 labels (wide-field detector) into prop+1.  The engine builds its
    //    candidate list and its detector medium from -1 / -2, so they are restored on a copy.
    //  * media: mua and mus were multiplied by cfg->unitinmm (src/mmc_mesh.c:542-546, :2396-2399); the engine applies the length unit
    //    itself (mmcb_config.unitinmm), so the copy divides it out again.
    std::vector<int> type(mesh->type, mesh->type + mesh->ne);

    for (int i = 0; i < mesh->srcelemlen; i++) {
        type[mesh->srcelem[i] - 1] = -1;
    }

    for (int i = 0; i < mesh->detelemlen; i++) {
        type[mesh->detelem[i] - 1] = -2;
    }

    std::vector<mmcb_medium> med(mesh->prop + 1);
    const float unit = (cfg->unitinmm > 0.f) ? cfg->unitinmm : 1.f;

    for (int i = 0; i <= mesh->prop; i++) {
        med[i].mua = mesh->med[i].mua / (i ? unit : 1.f);
        med[i].mus = mesh->med[i].mus / (i ? unit : 1.f);
        med[i].g = mesh->med[i].g;
        med[i].n = mesh->med[i].n;
    }

    mmcb_mesh m;
    memset(&m, 0, sizeof(m));
    m.nn = mesh->nn;
    m.ne = mesh->ne;
    m.prop = mesh->prop;
    m.node = node.data();
    m.elem = elem.data();
    m.type = type.data();
    m.med = med.data();
    m.facenb = facenb.empty() ? NULL : facenb.data();
    m.evol = mesh->evol;
    m.nvol = NULL;              // mesh->nvol already carries the surface correction of tracer_prep; the engine recomputes both

    // ---- configuration
    mmcb_config c;
    memset(&c, 0, sizeof(c));
    c.nphoton = cfg->nphoton;
    c.seed = cfg->seed;
    memcpy(c.srcpos, &cfg->srcpos, sizeof(c.srcpos));
    memcpy(c.srcdir, &cfg->srcdir, sizeof(c.srcdir));
    c.srctype = cfg->srctype;
    memcpy(c.srcparam1, &cfg->srcparam1, sizeof(c.srcparam1));
    memcpy(c.srcparam2, &cfg->srcparam2, sizeof(c.srcparam2));
    c.srcpattern = cfg->srcpattern;
    c.srcnum = cfg->srcnum;
    c.tstart = cfg->tstart;
    c.tstep = cfg->tstep;
    c.tend = cfg->tend;
    c.e0 = cfg->e0;
    c.isreflect = cfg->isreflect;
    c.isnormalized = cfg->isnormalized;
    c.issavedet = cfg->issavedet;
    c.ismomentum = cfg->ismomentum;
    c.issaveexit = cfg->issaveexit;
    c.issaveseed = cfg->issaveseed;
    c.isspecular = cfg->isspecular;
    c.issaveref = cfg->issaveref;
    // mcx_validatecfg coerces -M to a branch-less Badouel tracer for every GPU run (src/mmc_utils.c:3542-3544) before this function sees
    // cfg->method.  This engine offers Havel and Plucker on the GPU as well; the user's choice comes back in one of two ways:
    //   * with the one-line host patch of INTEGRATION.md (the coercion is skipped when this stub is linked) cfg->method IS the choice;
    //   * without any patch, MMC_B200_METHOD=p|h|s|g in the environment overrides the coerced value.
    c.method = cfg->method;

    if (const char* s = getenv("MMC_B200_METHOD")) {
        c.method = (s[0] == 'p') ? MMCB_RT_PLUCKER : (s[0] == 'h') ? MMCB_RT_HAVEL : (s[0] == 'g') ? MMCB_RT_BLBADOUEL_GRID : MMCB_RT_BLBADOUEL;
    }

    c.basisorder = cfg->basisorder;
    c.outputtype = cfg->outputtype;
    c.roulettesize = cfg->roulettesize;
    c.minenergy = cfg->minenergy;
    c.nout = cfg->nout;
    c.voidtime = cfg->voidtime;
    c.unitinmm = cfg->unitinmm;
    c.steps = cfg->steps.x;
    c.detnum = cfg->detnum;
    c.detpos = (const float*)cfg->detpos;
    c.maxdetphoton = cfg->maxdetphoton;
    c.photonseed = (const uint64_t*)cfg->photonseed;
    c.replayweight = cfg->replayweight;
    c.replaytime = cfg->replaytime;
    c.savetraj = (cfg->debuglevel & dlTraj) ? 1 : 0;
    c.maxjumpdebug = cfg->maxjumpdebug;
    c.nthread = cfg->autopilot ? 0 : cfg->nthread;
    c.nblocksize = cfg->autopilot ? 0 : cfg->nblocksize;
    c.respin = cfg->respin;
    // multi-slot sources (built by mcx_prep + mesh_init_srcdata_eid in the caller, src/mmc_host.c:136-165), RF, adjoint output
    c.omega = cfg->omega;
    c.srcid = cfg->srcid;
    c.extrasrclen = cfg->extrasrclen;
    c.srcdata = (const float*)cfg->srcdata;         // ExtraSrc = 4 x float4 (src/mmc_utils.h:147-152)
    c.detdir = (const float*)cfg->detdir;
    c.adjointmode = cfg->adjointmode;
    c.nodemua = (cfg->isnodalmua && cfg->nodemua) ? cfg->nodemua : NULL;          // src/mmc_cu_host.cu:477-487
    c.nodemusp = (c.nodemua && cfg->isnodalmusp && cfg->nodemusp) ? cfg->nodemusp : NULL;

    // devices: mcx_list_cu_gpu rewrote cfg->deviceid to the 1-based ids of the enabled GPUs (src/mmc_cu_host.cu:136-142).  One GPU: one
    // session (the mesh is prepared once, the output arrays are sized from it).  Several GPUs (-G 1101, -W a,b,c): the library shards
    // the photons by cfg->workload and merges the results over NCCL (mmcb_run_multi), which is the reference's omp fan-out (:1538-1553)
    std::vector<int> devices;
    std::vector<float> workload;

    for (unsigned int i = 0; i < activedev && cfg->deviceid[i] > 0; i++) {
        devices.push_back(cfg->deviceid[i] - 1);
        workload.push_back(cfg->workload[i]);
    }

    if (devices.empty()) {
        devices.push_back(0);
        workload.push_back(1.f);
    }

    float fullload = 0.f;

    for (float w : workload) {
        fullload += w;
    }

    if (fullload < 1e-6f) {         // unspecified: proportional to the core count (:407-412) -- equal on one box of identical GPUs
        for (size_t i = 0; i < workload.size(); i++) {
            workload[i] = (float)gpuinfo[devices[i]].core;
        }
    }

    const bool multi = devices.size() > 1;
    mmcb_session* sess = NULL;
    mmcb_sizes sz;

    if (multi) {
        B200_ASSERT(mmcb_query_sizes(&c, &m, &sz));
    } else {
        sess = mmcb_create(&c, &m, devices[0]);

        if (!sess) {
            B200_ASSERT(-1);
        }

        B200_ASSERT(mmcb_get_sizes(sess, &sz));
    }

    // ---- outputs (ownership as in src/mmc_cu_host.cu:339-367: exportfield defaults to mesh->weight; detected rows and
    //      seeds are malloc'ed here and freed by the caller)
#ifndef MCX_CONTAINER

    // The stock command-line host sizes mesh->weight for cfg->srcnum sources when the mesh is loaded (src/mmc_mesh.c:654,688) and
    // appends the detector / multi-source slots only later (mcx_prep, src/mmc_utils.c:3760-3797), so an adjoint run of the stock
    // program writes past the buffer.  Give the volume its real size here, as mesh_validate does for the containers
    // (src/mmc_mesh.c:2389-2394).
    if (cfg->exportfield == NULL && cfg->parentid == mpStandalone && sz.nslots > 1) {
        free(mesh->weight);
        mesh->weight = (double*)calloc(sz.fieldlen, sizeof(double));
    }

#endif

    if (cfg->exportfield == NULL) {
        cfg->exportfield = mesh->weight;
    }

    mmcb_output out;
    memset(&out, 0, sizeof(out));
    out.field = cfg->exportfield;
    out.dref = (cfg->issaveref) ? mesh->dref : NULL;

    if (cfg->issavedet) {
        cfg->exportdetected = (float*)realloc(cfg->exportdetected, sizeof(float) * (size_t)sz.reclen * cfg->maxdetphoton);
        out.detected = cfg->exportdetected;

        if (cfg->issaveseed) {
            cfg->exportseed = (unsigned char*)realloc(cfg->exportseed, 16 * (size_t)cfg->maxdetphoton);
            out.detseed = (uint64_t*)cfg->exportseed;
        }
    }

    if (c.savetraj) {
        cfg->exportdebugdata = (float*)realloc(cfg->exportdebugdata, sizeof(float) * 6 * (size_t)cfg->maxjumpdebug);
        out.traj = cfg->exportdebugdata;
    }

    // RF imaginary fluence (cfg->exportadjoint, float) and adjoint Jacobian (cfg->exportjacob), src/mmc_cu_host.cu:345-353,1203-1242
    const bool isrf = (cfg->omega > 0.f && cfg->seed != SEED_FROM_FILE);
    std::vector<double> field_im;

    if (isrf) {
        field_im.assign(sz.fieldlen, 0.0);
        out.field_im = field_im.data();

        if (cfg->exportadjoint == NULL) {
            cfg->exportadjoint = (float*)calloc(sz.fieldlen, sizeof(float));
        }
    }

    if (sz.jacoblen) {
        if (cfg->exportjacob) {
            free(cfg->exportjacob);
        }

        cfg->exportjacob = (float*)calloc(sz.jacoblen, sizeof(float));
        out.jacob = cfg->exportjacob;
    }

    static const char* tracername[] = {"Plucker", "Havel", "Badouel", "branch-less Badouel", "branch-less Badouel + dual grid"};
    MMC_FPRINTF(cfg->flog, "- code name: [MMC-B200] sm_100a photon engine (libmmc_b200 %x), tracer: %s\n", mmcb_version(),
                tracername[(c.method >= 0 && c.method <= 4) ? c.method : 3]);
    {
        std::vector<uint64_t> share(devices.size());
        mmcb_photon_shares((uint64_t)cfg->nphoton, (int)devices.size(), workload.data(), share.data());

        for (size_t i = 0; i < devices.size(); i++) {
            MMC_FPRINTF(cfg->flog, "- [device %d(%d): %s] np=%.1f maxgate=%d repetition=%d\n", gpuinfo[devices[i]].id, (int)i + 1, gpuinfo[devices[i]].name,
                        (double)share[i], sz.maxgate, cfg->respin);
        }
    }

    MMC_FPRINTF(cfg->flog, "lauching mmc_main_loop for time window [%.1fns %.1fns] ...\n", cfg->tstart * 1e9, cfg->tend * 1e9);
    mcx_fflush(cfg->flog);
    unsigned int tic = StartTimer();

    if (multi) {
        B200_ASSERT(mmcb_run_multi(&c, &m, (int)devices.size(), devices.data(), workload.data(), &out));
    } else {
        B200_ASSERT(mmcb_run_session(sess, &out));
        mmcb_destroy(sess);
    }

    unsigned int toc = GetTimeMillis() - tic;
    MMC_FPRINTF(cfg->flog, "kernel complete:  \t%d ms\nretrieving flux ... \t", (int)(out.kernel_ms + 0.5f));
    MMC_FPRINTF(cfg->flog, "transfer complete:        %d ms\n", toc);

    for (size_t i = 0; i < field_im.size(); i++) {
        cfg->exportadjoint[i] += (float)field_im[i];
    }

    // ---- scalar results the callers read (src/mmc_cu_host.cu:757-759,775-782,823-853,995)
    cfg->runtime = (unsigned int)(out.kernel_ms + 0.5f);
    cfg->his.normalizer = (float)out.normalizer;
    cfg->normalizer = (float)out.normalizer;
    cfg->detectedcount = out.detectedcount;
    cfg->his.detected = out.detectedtotal;
    cfg->debugdatalen = out.trajcount;

    if (cfg->issavedet && out.detectedtotal > 0) {      // src/mmc_cu_host.cu:828
        MMC_FPRINTF(cfg->flog, "detected %d photons, total: %d\t", (int)out.detectedcount, (int)out.detectedtotal);
    }

    if (out.detectedtotal > cfg->maxdetphoton) {
        MMC_FPRINTF(cfg->flog, S_RED "WARNING: the detected photon (%d) is more than what your have specified (%d), please use the -H option to specify a greater number\t" S_RESET,
                    out.detectedtotal, cfg->maxdetphoton);
    }

    double energytot = 0., energyesc = 0.;

    for (int j = 0; j < cfg->srcnum; j++) {
        energytot += out.energytot[j];
        energyesc += out.energyesc[j];
    }

    // ---- files, exactly the calls of src/mmc_cu_host.cu:1399-1438
#ifndef MCX_CONTAINER

    if (cfg->issave2pt && cfg->parentid == mpStandalone) {
        MMC_FPRINTF(cfg->flog, "saving data to file ...\t");
        mesh_saveweight(mesh, cfg, 0);
        MMC_FPRINTF(cfg->flog, "saving data complete : %d ms\n\n", GetTimeMillis() - tic);
    }

    if (cfg->issavedet && cfg->parentid == mpStandalone && cfg->exportdetected) {
        cfg->his.totalphoton = cfg->nphoton;
        cfg->his.unitinmm = cfg->unitinmm;
        cfg->his.savedphoton = cfg->detectedcount;
        cfg->his.colcount = sz.reclen;
        cfg->his.seedbyte = (cfg->exportseed) ? 16 : 0;
        mcx_savedetphoton(cfg->exportdetected, (void*)(cfg->exportseed), cfg->detectedcount, 0, cfg);
    }

    if (c.savetraj && cfg->parentid == mpStandalone && cfg->exportdebugdata) {
        cfg->his.colcount = 6;
        cfg->his.savedphoton = cfg->debugdatalen;
        cfg->his.totalphoton = cfg->nphoton;
        cfg->his.detected = 0;
        mcx_savedetphoton(cfg->exportdebugdata, NULL, cfg->debugdatalen, 0, cfg);
    }

    if (cfg->issaveref) {
        mesh_saveweight(mesh, cfg, 1);
    }

    if (cfg->issave2pt && cfg->parentid == mpStandalone && cfg->exportjacob && sz.jacoblen) {     // src/mmc_cu_host.cu:1244-1250,1385-1391
        mesh_savejacob(cfg, mesh, cfg->exportjacob, sz.adj_ns, sz.adj_nd, isrf ? 1 : 0, MCX_IS_DUAL_ADJOINT_TYPE(cfg->outputtype));
    }

#endif
    // the two lines scripts parse (src/mmc_cu_host.cu:1444-1453)
    MMC_FPRINTF(cfg->flog, "simulated %ld photons (%ld) with devices (ray-tet %.0f)\nMCX simulation speed: %.2f photon/ms\n",
                (long)cfg->nphoton, (long)cfg->nphoton, out.raytet, (double)cfg->nphoton / (out.kernel_ms > 0.f ? out.kernel_ms : 1.f));
    MMC_FPRINTF(cfg->flog, "total simulated energy: %.2f\tabsorbed: %5.5f%%\n(loss due to initial specular reflection is excluded in the total)\n",
                energytot, (energytot - energyesc) / energytot * 100.f);
    mcx_fflush(cfg->flog);
    free(gpuinfo);
}
