// glue_driver.C -- TEST SCAFFOLDING: a stand-in "solver" that includes the two glue snippets
// exactly the way cudaParticlesUncoupledFoam.C / cudaParticlesPimpleFoam.C do, over the foam shim.
//   glue_driver <case.bin> <out.bin> <nEulerianSteps>
// case.bin (written by tests/test_gpu_glue.py): header of int32 {nPoints,nFaces,nInternal,nCells,
// nPatches,numParticles,saveInterval,randomWalk(0 none,1 xorwow,2 philox)}, doubles {deltaT, dt, D,
// seedLo[3], seedHi[3]}, then points, faceOffsets, faceVerts, owner, neighbour, centres,
// patchStart[nPatches+1], and nEulerianSteps cell fields U[nCells][3].
#include <cstring>
#include <fstream>

#include "foam_shim.H"
#include "cpf.h"

using namespace Foam;

// Decomposed runs: start one process per rank with CPF_SHIM_NPROCS / CPF_SHIM_RANK / CPF_SHIM_DIR set and give every
// rank the case file of ITS processor mesh (tests/test_glue.py splits the global mesh); the shim's Pstream carries the
// gathers through files, the field and the statistics travel over NCCL inside libcpf.

template <class T> static void rd(std::ifstream &f, T *p, size_t n) { f.read(reinterpret_cast<char *>(p), sizeof(T) * n); }

int main(int argc, char **argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: glue_driver case.bin out.bin nSteps\n"); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    int h[8];
    double d[9];
    rd(in, h, 8);
    rd(in, d, 9);
    const int nPoints = h[0], nFaces = h[1], nInternal = h[2], nCells = h[3], nPatches = h[4];
    fvMesh mesh;
    mesh.points_.setSize(nPoints); rd(in, mesh.points_.data(), nPoints);
    std::vector<int> off(nFaces + 1); rd(in, off.data(), nFaces + 1);
    std::vector<int> fv(off[nFaces]); rd(in, fv.data(), fv.size());
    mesh.faces_.setSize(nFaces);
    for (int f = 0; f < nFaces; ++f) { mesh.faces_[f].setSize(off[f + 1] - off[f]); for (int k = off[f]; k < off[f + 1]; ++k) mesh.faces_[f][k - off[f]] = fv[k]; }
    mesh.owner_.setSize(nFaces); rd(in, mesh.owner_.data(), nFaces);
    mesh.neighbour_.setSize(nInternal); rd(in, mesh.neighbour_.data(), nInternal);
    mesh.C_.setSize(nCells); rd(in, mesh.C_.data(), nCells);
    std::vector<int> ps(nPatches + 1); rd(in, ps.data(), nPatches + 1);
    mesh.patches_.setSize(nPatches);
    for (int p = 0; p < nPatches; ++p) mesh.patches_[p] = polyPatch{"patch" + std::to_string(p), ps[p], ps[p + 1] - ps[p]};
    mesh.tetBase_ = labelList(nFaces, 0);
    mesh.cells_.setSize(nCells);
    for (int f = 0; f < nFaces; ++f) {
        mesh.cells_[mesh.owner_[f]].v.push_back(f);
        if (f < nInternal) mesh.cells_[mesh.neighbour_[f]].v.push_back(f);
    }

    IOdictionary cudaParticleAdvectionDict;
    cudaParticleAdvectionDict.kv["numParticles"] = std::to_string(h[5]);
    cudaParticleAdvectionDict.kv["saveInterval"] = std::to_string(h[6]);
    cudaParticleAdvectionDict.kv["randomWalk"] = h[7] == 0 ? "none" : (h[7] == 1 ? "xorwow" : "philox");
    { std::ostringstream o; o.precision(17); o << d[1]; cudaParticleAdvectionDict.kv["dt"] = o.str(); }
    { std::ostringstream o; o.precision(17); o << d[2]; cudaParticleAdvectionDict.kv["diffusionCoeff"] = o.str(); }
    cudaParticleAdvectionDict.boxes.emplace("seedingBox", boundBox(point(d[3], d[4], d[5]), point(d[6], d[7], d[8])));
    if (const char *extra = std::getenv("CPF_SHIM_DICT")) { // "key=value;key=value": further dictionary entries
        std::istringstream is(extra);
        std::string kv;
        while (std::getline(is, kv, ';')) {
            const size_t eq = kv.find('=');
            if (eq != std::string::npos) cudaParticleAdvectionDict.kv[kv.substr(0, eq)] = kv.substr(eq + 1);
        }
    }
    if (h[7] < 0) cudaParticleAdvectionDict.kv.erase("randomWalk"); // the stock dictionary: no optional key at all

    Time runTime;
    runTime.dT = d[0];
    volVectorField U;
    U.f.setSize(nCells);
    const int nSteps = std::atoi(argv[3]);
    rd(in, U.f.data(), nCells);

    #include "initCuda.H"

    for (int it = 0; it < nSteps; ++it)   // the pimpleFoam-style time loop (cudaParticlesPimpleFoam.C:130-195)
    {
        runTime.t += runTime.dT;
        if (it > 0) rd(in, U.f.data(), nCells);   // "the flow solve"
        #include "advect.H"
    }

    const long long n = cpf_num_particles(cpf);
    std::vector<double> p(4 * n), v(4 * n);
    std::vector<int> tet(n);
    CPF_GLUE_CHECK(cpf_download(cpf, p.data(), v.data(), tet.data()));
    std::ofstream out(argv[2], std::ios::binary);
    out.write(reinterpret_cast<const char *>(&n), sizeof n);
    out.write(reinterpret_cast<const char *>(p.data()), sizeof(double) * p.size());
    out.write(reinterpret_cast<const char *>(v.data()), sizeof(double) * v.size());
    out.write(reinterpret_cast<const char *>(tet.data()), sizeof(int) * tet.size());
    out.write(reinterpret_cast<const char *>(&step), sizeof step);
    std::vector<int> cell(n);   // global cell ids: comparable between a serial run and a decomposed one
    CPF_GLUE_CHECK(cpf_download_cells(cpf, cell.data()));
    out.write(reinterpret_cast<const char *>(cell.data()), sizeof(int) * cell.size());
    cpf_destroy(cpf);
    return 0;
}
