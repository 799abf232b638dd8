// Host-side Mesh: loading, merging and transforming geometry before it is handed to the GPU.
// Behavioural mirror of the reference's mesh.cpp (file:line cited per routine); written from
// scratch around a memory-buffer tokenizer instead of fscanf.
#include "mesh.h"

#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>

#include "trace.h"

namespace
{

// Cornell box, 16 quads, original Cornell units (www.graphics.cornell.edu/online/box/data.html);
// same quad order and vertex order as the reference's table (cornell_box.cpp:7-88), stored flat
const float kLightY = 548.8f - 0.1f; // just below the ceiling, as in the reference (cornell_box.cpp:39-43)
const float kCornell[16][12] = {
    { 552.8f, 0, 0,  0, 0, 0,  0, 0, 559.2f,  549.6f, 0, 559.2f },                       // floor
    { 556, 548.8f, 0,  556, 548.8f, 559.2f,  0, 548.8f, 559.2f,  0, 548.8f, 0 },         // ceiling
    { 549.6f, 0, 559.2f,  0, 0, 559.2f,  0, 548.8f, 559.2f,  556, 548.8f, 559.2f },      // back wall
    { 0, 0, 559.2f,  0, 0, 0,  0, 548.8f, 0,  0, 548.8f, 559.2f },                       // right wall
    { 552.8f, 0, 0,  549.6f, 0, 559.2f,  556, 548.8f, 559.2f,  556, 548.8f, 0 },         // left wall
    { 343, kLightY, 227,  343, kLightY, 332,  213, kLightY, 332,  213, kLightY, 227 },   // light
    { 130, 165, 65,  82, 165, 225,  240, 165, 272,  290, 165, 114 },                     // short block
    { 290, 0, 114,  290, 165, 114,  240, 165, 272,  240, 0, 272 },
    { 130, 0, 65,  130, 165, 65,  290, 165, 114,  290, 0, 114 },
    { 82, 0, 225,  82, 165, 225,  130, 165, 65,  130, 0, 65 },
    { 240, 0, 272,  240, 165, 272,  82, 165, 225,  82, 0, 225 },
    { 423, 330, 247,  265, 330, 296,  314, 330, 456,  472, 330, 406 },                   // tall block
    { 423, 0, 247,  423, 330, 247,  472, 330, 406,  472, 0, 406 },
    { 472, 0, 406,  472, 330, 406,  314, 330, 456,  314, 0, 456 },
    { 314, 0, 456,  314, 330, 456,  265, 330, 296,  265, 0, 296 },
    { 265, 0, 296,  265, 330, 296,  423, 330, 247,  423, 0, 247 },
};

// Whitespace separated tokens of a text buffer
struct Tokens
{
    const char *p, *end;

    void skip_space()
    {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n' || *p == '\f' || *p == '\v'))
            p++;
    }
    bool at_end()
    {
        skip_space();
        return p >= end;
    }
    // "%f" semantics = strtof.  Plain decimals take a fast exact path (Clinger): up to 2^53 as an integer
    // mantissa times / over a power of ten <= 10^22 is a correctly rounded double, and narrowing that to float
    // rounds like the decimal itself unless the double sits exactly on a float rounding midpoint (then, and for
    // anything unusual -- inf, nan, hex floats, subnormal or huge values, very long mantissas -- strtof decides)
    bool next_float(float& out)
    {
        static const double kPow10[23] = { 1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22 };
        skip_space();
        if (p >= end)
            return false;
        const char *s = p;
        bool neg = false;
        if (s < end && (*s == '-' || *s == '+'))
            neg = *s++ == '-';
        unsigned long long mant = 0;
        int digits = 0, exp10 = 0;
        bool any = false, fast = true;
        for (; s < end && *s >= '0' && *s <= '9'; s++, any = true)
        {
            if (mant || *s != '0')
                digits++;
            if (digits <= 18)
                mant = mant * 10 + (unsigned) (*s - '0');
            else
                fast = false;
        }
        if (s < end && *s == '.')
        {
            s++;
            for (; s < end && *s >= '0' && *s <= '9'; s++, any = true)
            {
                if (mant || *s != '0')
                    digits++;
                if (digits <= 18)
                {
                    mant = mant * 10 + (unsigned) (*s - '0');
                    exp10--;
                }
                else
                    fast = false;
            }
        }
        if (any && fast && s < end && (*s == 'e' || *s == 'E'))
        {
            const char *e = s + 1;
            bool eneg = false;
            if (e < end && (*e == '-' || *e == '+'))
                eneg = *e++ == '-';
            if (e < end && *e >= '0' && *e <= '9')
            {
                int ev = 0;
                for (; e < end && *e >= '0' && *e <= '9'; e++)
                    ev = ev < 10000 ? ev * 10 + (*e - '0') : ev;
                exp10 += eneg ? -ev : ev;
                s = e;
            }
        }
        // what follows must end the number the way strtof would see it (hex floats, "inf", "nan" start differently)
        const bool plain_end = s >= end || *s == ' ' || *s == '\t' || *s == '\r' || *s == '\n' || *s == '\f' || *s == '\v';
        if (any && fast && plain_end && mant < (1ull << 53) && exp10 >= -22 && exp10 <= 22)
        {
            const double v = exp10 < 0 ? (double) mant / kPow10[-exp10] : (double) mant * kPow10[exp10];
            unsigned long long bits;
            std::memcpy(&bits, &v, sizeof(bits));
            const bool on_midpoint = (bits & 0x1FFFFFFFull) == 0x10000000ull;
            if (!on_midpoint && (v == 0.0 || (v > 1.2e-38 && v < 3.4e38)))
            {
                const float f = (float) v;
                out = neg ? -f : f;
                p = s;
                return true;
            }
        }
        char *stop = nullptr;
        out = std::strtof(p, &stop);
        if (stop == p)
            return false;
        p = stop;
        return true;
    }
    // "%i" semantics: decimal, or 0x.. / 0.. prefixes, optional sign
    bool next_int(long& out)
    {
        skip_space();
        if (p >= end)
            return false;
        char *stop = nullptr;
        out = std::strtol(p, &stop, 0);
        if (stop == p)
            return false;
        p = stop;
        return true;
    }
};

bool fail(const char *msg)
{
    Trace("Mesh::Read() - %s", msg);
    return false;
}

int count_floats_on_line(const char *line, const char *line_end)
{
    int n = 0;
    Tokens t = { line, line_end };
    float f;
    while (n < 9 && t.next_float(f))
        n++;
    return n;
}

} // namespace

Vec3f TriangleNormal(const Vec3f& v0, const Vec3f& v1, const Vec3f& v2)
{
    return Normalize(Cross(v1 - v0, v2 - v0)); // reference triangle.h:109-114
}

void Mesh::Clear()
{
    m_triangles.clear();
    m_vertices.clear();
}

void Mesh::CornellBox() // reference mesh.cpp:18-24
{
    Clear();
    for (const auto& quad : kCornell)
        AddQuad(quad);
}

// Two triangles (0,1,2) (0,2,3) sharing four new vertices that all carry the quad's face normal
// (reference mesh.cpp:26-52)
void Mesh::AddQuad(const float *q)
{
    const Vec3f corner[4] = { Vec3f(q), Vec3f(q + 3), Vec3f(q + 6), Vec3f(q + 9) };
    const Vec3f n = TriangleNormal(corner[0], corner[1], corner[2]);
    const uint32 first = uint32(m_vertices.size());
    for (const Vec3f& c : corner)
        m_vertices.push_back(Vertex { c, n });
    m_triangles.push_back(Triangle { first, first + 1, first + 2, n });
    m_triangles.push_back(Triangle { first, first + 2, first + 3, n });
}

void Mesh::AddMesh(const Mesh& other) // reference mesh.cpp:54-70
{
    const uint32 shift = uint32(m_vertices.size());
    m_vertices.insert(m_vertices.end(), other.m_vertices.begin(), other.m_vertices.end());
    m_triangles.reserve(m_triangles.size() + other.m_triangles.size());
    for (Triangle t : other.m_triangles)
    {
        t.v0 += shift;
        t.v1 += shift;
        t.v2 += shift;
        m_triangles.push_back(t);
    }
}

// Box of the vertices that triangles actually reference.  Like the reference (mesh.cpp:72-94)
// the maximum starts from numeric_limits<float>::min() -- the smallest POSITIVE float -- so an
// all-negative axis reports a maximum of ~0; downstream results (grid box) depend on it.
void Mesh::ComputeAABB(Vec3f& aabb_min, Vec3f& aabb_max) const
{
    if (m_triangles.empty())
    {
        aabb_min = aabb_max = Vec3f(0.0f);
        return;
    }
    aabb_min = Vec3f(std::numeric_limits<float>::max());
    aabb_max = Vec3f(std::numeric_limits<float>::min());
    for (const Triangle& t : m_triangles)
        for (uint32 vi : { t.v0, t.v1, t.v2 })
        {
            aabb_min = ComponentMin(aabb_min, m_vertices[vi].p);
            aabb_max = ComponentMax(aabb_max, m_vertices[vi].p);
        }
}

// Positions by mat, normals by its inverse transpose, renormalised (reference mesh.cpp:96-118).
// A singular matrix leaves the normals transformed by the untouched matrix' transpose, as the
// reference does once its assert is compiled out.
void Mesh::Transform(Matrix44f mat)
{
    Matrix44f normal_mat = mat;
    normal_mat.Invert();
    normal_mat.Transpose4x4();
    for (Triangle& t : m_triangles)
    {
        normal_mat.Transf3x3(t.n);
        t.n = Normalize(t.n);
    }
    for (Vertex& v : m_vertices)
    {
        mat.Transf4x4(v.p);
        normal_mat.Transf3x3(v.n);
        v.n = Normalize(v.n);
    }
}

// Centre on the origin and scale the longest side to 1 (reference mesh.cpp:120-136)
void Mesh::NormalizeDimensions()
{
    Vec3f lo, hi;
    ComputeAABB(lo, hi);
    const Vec3f centre = (lo + hi) / 2.0f;
    const Vec3f size = hi - lo;
    Matrix44f shift, scale;
    shift.Translation(-centre.x, -centre.y, -centre.z);
    scale.Scaling(1.0f / std::max(std::max(size.x, size.y), size.z));
    Transform(shift * scale);
}

void Mesh::SetArrays(const float *vtx6, uint32 num_vtx, const uint32 *tri6, uint32 num_tri)
{
    m_vertices.resize(num_vtx);
    m_triangles.resize(num_tri);
    if (num_vtx)
        std::memcpy(static_cast<void *>(m_vertices.data()), vtx6, size_t(num_vtx) * sizeof(Vertex));
    if (num_tri)
        std::memcpy(static_cast<void *>(m_triangles.data()), tri6, size_t(num_tri) * sizeof(Triangle));
}

bool Mesh::ReadBinary(const char *filename)
{
    Clear();
    std::FILE *f = std::fopen(filename, "rb");
    if (!f)
        return fail("Can't open file");
    char magic[8];
    uint32 counts[2];
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "RTMMESH1", 8) == 0 &&
              std::fread(counts, 4, 2, f) == 2;
    if (ok)
    {
        m_vertices.resize(counts[0]);
        m_triangles.resize(counts[1]);
        ok = std::fread(static_cast<void *>(m_vertices.data()), sizeof(Vertex), counts[0], f) == counts[0] &&
             std::fread(static_cast<void *>(m_triangles.data()), sizeof(Triangle), counts[1], f) == counts[1];
    }
    std::fclose(f);
    if (ok)
        for (const Triangle& t : m_triangles)
            ok = ok && t.v0 < counts[0] && t.v1 < counts[0] && t.v2 < counts[0];
    if (!ok)
    {
        Clear();
        return fail("Bad binary mesh");
    }
    return true;
}

// ASCII ".dat" meshes (reference mesh.cpp:138-391).  Two layouts:
//   non-indexed: one vertex per line, three consecutive vertices form a triangle
//   indexed:     <num vertices> <vertices...> <num indices> <indices...>
// A file is indexed iff its first line contains no space.  A vertex line carries
//   3 floats (position) | 6 (+normal) | 8 (+uv, ignored) | 9 (+rgb, ignored)
// decided from the first vertex line.  Without normals every vertex gets the face normal of the
// last triangle that references it.  flip_winding swaps v0/v1 before the face normal is taken.
bool Mesh::Read(const char *filename, bool flip_winding)
{
    Clear();

    std::FILE *f = std::fopen(filename, "rb");
    if (!f)
        return fail("Can't open file");
    std::string text;
    char chunk[1 << 16];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0)
        text.append(chunk, got);
    std::fclose(f);
    if (text.empty())
        return fail("Can't read 1st line");

    const char *begin = text.data(), *end = begin + text.size();
    const char *eol = begin;
    while (eol < end && *eol != '\n')
        eol++;
    bool indexed = true;
    for (const char *c = begin; c < eol; c++)
        if (*c == ' ')
            indexed = false;

    Tokens tok = { begin, end };
    long num_vtx = 0;
    if (indexed)
    {
        if (!tok.next_int(num_vtx))
            return fail("Can't get vertex count");
        if (num_vtx < 3)
            return fail("Invalid vertex count");
    }

    // vertex layout from the first vertex line
    tok.skip_space();
    const char *line_end = tok.p;
    while (line_end < end && *line_end != '\n')
        line_end++;
    if (tok.p >= end)
        return fail("Can't read 1st vertex");
    const int per_line = count_floats_on_line(tok.p, line_end);
    if (per_line != 3 && per_line != 6 && per_line != 8 && per_line != 9)
        return fail("Invalid vertex spec");
    const int extras = per_line > 6 ? per_line - 6 : 0;

    // vertices
    uint vtx_read = 0;
    while (!tok.at_end())
    {
        Vertex v;
        v.n = Vec3f(0.0f);
        if (!tok.next_float(v.p.x) || !tok.next_float(v.p.y) || !tok.next_float(v.p.z))
            return fail("Can't read position");
        if (per_line >= 6 && (!tok.next_float(v.n.x) || !tok.next_float(v.n.y) || !tok.next_float(v.n.z)))
            return fail("Can't read normal");
        for (int i = 0; i < extras; i++)
        {
            float ignored;
            if (!tok.next_float(ignored))
                return fail(extras == 2 ? "Can't read UV" : "Can't read RGB");
        }
        m_vertices.push_back(v);
        vtx_read++;
        if (indexed && vtx_read >= uint(num_vtx))
            break;
    }
    if (vtx_read == 0 || (indexed && vtx_read != uint(num_vtx)))
        return fail("Can't read all vertices");
    if (!indexed && vtx_read % 3 != 0)
        return fail("Invalid vertex count");

    // triangles
    uint num_tri;
    if (indexed)
    {
        long num_idx = 0;
        if (!tok.next_int(num_idx))
            return fail("Can't get index count");
        if (num_idx < 3 || num_idx % 3 != 0)
            return fail("Invalid index count");
        num_tri = uint(num_idx / 3);
    }
    else
        num_tri = vtx_read / 3;
    m_triangles.resize(num_tri);
    for (uint i = 0; i < num_tri; i++)
    {
        Triangle& t = m_triangles[i];
        if (indexed)
        {
            long a, b, c;
            if (!tok.next_int(a) || !tok.next_int(b) || !tok.next_int(c))
                return fail("Can't read triangle indices");
            t.v0 = uint32(a);
            t.v1 = uint32(b);
            t.v2 = uint32(c);
            if (t.v0 >= vtx_read || t.v1 >= vtx_read || t.v2 >= vtx_read)
                return fail("Vertex index out of bounds");
        }
        else
        {
            t.v0 = i * 3;
            t.v1 = i * 3 + 1;
            t.v2 = i * 3 + 2;
        }
        if (flip_winding)
            std::swap(t.v0, t.v1);
        t.n = TriangleNormal(m_vertices[t.v0].p, m_vertices[t.v1].p, m_vertices[t.v2].p);
    }

    if (per_line == 3)
        for (const Triangle& t : m_triangles)
            m_vertices[t.v0].n = m_vertices[t.v1].n = m_vertices[t.v2].n = t.n;

    Trace("Loaded mesh '%s', NumVtx: %i, NumTri: %i, %s, %s", filename, vtx_read, num_tri,
          indexed ? "Indexed" : "Non-Indexed",
          per_line == 3 ? "VSPos" : per_line == 6 ? "VSPosNormal" : per_line == 8 ? "VSPosNormalUV" : "VSPosNormalRGB");
    return true;
}
