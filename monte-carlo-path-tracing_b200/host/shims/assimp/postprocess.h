// assimp/postprocess.h — TEST INFRASTRUCTURE ONLY (oracle build shim).
// Only the flag values used at src/parser/model_loader.cpp:512-520 of the reference.
#ifndef ORACLE_SHIM_ASSIMP_POSTPROCESS_H
#define ORACLE_SHIM_ASSIMP_POSTPROCESS_H
enum aiPostProcessSteps {
    aiProcess_CalcTangentSpace = 0x1,
    aiProcess_Triangulate = 0x8,
    aiProcess_GenSmoothNormals = 0x40,
    aiProcess_GenUVCoords = 0x40000,
    aiProcess_FlipUVs = 0x800000
};
#endif
