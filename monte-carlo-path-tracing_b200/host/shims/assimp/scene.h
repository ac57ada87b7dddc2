// assimp/scene.h — TEST INFRASTRUCTURE ONLY (oracle build shim).
// The POD subset of assimp's scene graph that the reference's
// ProcessAssimpNode (src/parser/model_loader.cpp:335-419) reads.
#ifndef ORACLE_SHIM_ASSIMP_SCENE_H
#define ORACLE_SHIM_ASSIMP_SCENE_H

#define AI_SCENE_FLAGS_INCOMPLETE 0x1
#define AI_MAX_NUMBER_OF_TEXTURECOORDS 8

struct aiVector3D {
    float x, y, z;
};

struct aiFace {
    unsigned int mNumIndices;
    unsigned int *mIndices;
};

struct aiMesh {
    unsigned int mNumVertices;
    unsigned int mNumFaces;
    aiVector3D *mVertices;
    aiVector3D *mNormals;
    aiVector3D *mTangents;
    aiVector3D *mBitangents;
    aiVector3D *mTextureCoords[AI_MAX_NUMBER_OF_TEXTURECOORDS];
    aiFace *mFaces;
};

struct aiNode {
    unsigned int mNumMeshes;
    unsigned int *mMeshes;
    unsigned int mNumChildren;
    aiNode **mChildren;
};

struct aiScene {
    unsigned int mFlags;
    aiNode *mRootNode;
    unsigned int mNumMeshes;
    aiMesh **mMeshes;
};

#endif
