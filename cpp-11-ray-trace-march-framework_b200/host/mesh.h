// Triangle mesh container with the reference's public layout and methods (reference
// mesh.h:10-37): 24-byte Triangle {v0,v1,v2,n} and Vertex {p,n} records in two public vectors,
// which is exactly what cuda_trace_upload_scene consumes.
#ifndef RTM_HOST_MESH_H
#define RTM_HOST_MESH_H

#include <vector>

#include "lin_alg.h"
#include "types.h"

struct Mesh
{
    struct Triangle
    {
        uint32 v0, v1, v2;
        Vec3f  n; // face normal
    };

    struct Vertex
    {
        Vec3f p; // position
        Vec3f n; // shading normal
    };

    std::vector<Triangle> m_triangles;
    std::vector<Vertex>   m_vertices;

    // -- sources: an ASCII .dat file (position / +normal / +uv per line, indexed or not; false + a Trace line on
    //    failure), the built-in Cornell box (16 quads), or nothing
    bool Read(const char *filename, bool flip_winding = false);
    void CornellBox();
    void Clear();

    // -- composition: one quad as two triangles (4 x 3 floats, flat normal), or all of another mesh
    void AddQuad(const float *quad_vtx);
    void AddMesh(const Mesh& mesh);

    // -- geometry: bounds over the referenced vertices, row-vector transform of positions and normals, and
    //    "centre at the origin, longest side = 1"
    void ComputeAABB(Vec3f& aabb_min, Vec3f& aabb_max) const;
    void Transform(Matrix44f mat);
    void NormalizeDimensions();

    // Additions (not in the reference): the binary asset format of oracle/convert_meshes.py --
    // the state of a Mesh right after Read() -- and direct array access for the C wrappers
    bool ReadBinary(const char *filename);
    void SetArrays(const float *vtx6, uint32 num_vtx, const uint32 *tri6, uint32 num_tri);
};

static_assert(sizeof(Mesh::Triangle) == 24 && sizeof(Mesh::Vertex) == 24, "records must stay 24 bytes");

Vec3f TriangleNormal(const Vec3f& v0, const Vec3f& v1, const Vec3f& v2);

#endif
