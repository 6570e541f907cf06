// The compact read-back's wire format (MLSB_READBACK=compact): shared by marching.cu (compactMeshKernel) and
// unpack.cpp (plain C++, AVX2 + FMA where the CPU has them), which turns the wire blob of a block into its
// HostKeyMesh blob.
#ifndef MLSB200_WIRE_H
#define MLSB200_WIRE_H

#include <cstdint>

namespace mlsb
{

struct WireParams
{
    uint64_t numVertices, numTriangles, numInternalVertices;
    uint32_t zFirst;          // first corner slice of the block (the lattice z of a vertex is relative to it)
    uint32_t gridOffset[3];   // Marching::generate's keyOffset
    float scale, bias[3];     // ScaleBiasFilter
};

void unpackWireMesh(const uint64_t *wire, const WireParams &p, void *hostBlob);

}  // namespace mlsb

#endif
