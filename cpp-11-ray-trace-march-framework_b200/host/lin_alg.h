// Host-side vector / matrix types with the reference's public spellings (Vec2f, Vec3f, Vec2ui,
// Matrix44f, Dot, Cross, Normalize, ComponentMin/Max, DegToRad, ToBGRA8, Clamp) -- the subset the
// Mesh / Scene / Renderer API surface needs (reference lin_alg.h:17-232, 235-690).
//
// Only behaviour is mirrored, not text: every routine evaluates its fp32 expression in the same
// operand order as the reference (products summed left to right, no FMA: host objects are built
// with -ffp-contract=off), because mesh transforms feed vertex positions whose bits decide
// ray/triangle hits.  tests/test_host_vs_ref.py compares every routine with the reference's.
#ifndef RTM_HOST_LIN_ALG_H
#define RTM_HOST_LIN_ALG_H

#include <algorithm>
#include <cmath>
#include <cstring>

#include "types.h"

struct Vec2f
{
    float x, y;
    Vec2f() : x(0.0f), y(0.0f) { }
    Vec2f(float x_, float y_) : x(x_), y(y_) { }
};

struct Vec2ui
{
    uint x, y;
    Vec2ui() : x(0), y(0) { }
    Vec2ui(uint x_, uint y_) : x(x_), y(y_) { }
};

struct Vec3f
{
    float x, y, z;

    Vec3f() : x(0.0f), y(0.0f), z(0.0f) { }
    Vec3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) { }
    explicit Vec3f(float s) : x(s), y(s), z(s) { }
    explicit Vec3f(const float *p) : x(p[0]), y(p[1]), z(p[2]) { }

    float  operator [] (uint i) const { return (&x)[i]; }
    float& operator [] (uint i)       { return (&x)[i]; }

    Vec3f operator + (const Vec3f& o) const { return Vec3f(x + o.x, y + o.y, z + o.z); }
    Vec3f operator - (const Vec3f& o) const { return Vec3f(x - o.x, y - o.y, z - o.z); }
    Vec3f operator * (const Vec3f& o) const { return Vec3f(x * o.x, y * o.y, z * o.z); }
    Vec3f operator / (const Vec3f& o) const { return Vec3f(x / o.x, y / o.y, z / o.z); }
    Vec3f operator + (float s) const { return Vec3f(x + s, y + s, z + s); }
    Vec3f operator - (float s) const { return Vec3f(x - s, y - s, z - s); }
    Vec3f operator * (float s) const { return Vec3f(x * s, y * s, z * s); }
    Vec3f operator / (float s) const { return Vec3f(x / s, y / s, z / s); }
    Vec3f operator - () const { return Vec3f(-x, -y, -z); }
    void operator += (const Vec3f& o) { x += o.x; y += o.y; z += o.z; }
    void operator -= (const Vec3f& o) { x -= o.x; y -= o.y; z -= o.z; }
    void operator *= (float s) { x *= s; y *= s; z *= s; }
    void operator /= (float s) { x /= s; y /= s; z /= s; }
    bool operator == (const Vec3f& o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator != (const Vec3f& o) const { return !(*this == o); }
};

inline Vec3f operator * (float s, const Vec3f& v) { return Vec3f(v.x * s, v.y * s, v.z * s); }

// Accumulates from zero, left to right (reference lin_alg.h:138-144)
inline float Dot(const Vec3f& a, const Vec3f& b)
{
    float acc = 0.0f;
    acc += a.x * b.x;
    acc += a.y * b.y;
    acc += a.z * b.z;
    return acc;
}

inline Vec3f Cross(const Vec3f& a, const Vec3f& b)
{
    return Vec3f(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline Vec3f operator ^ (const Vec3f& a, const Vec3f& b) { return Cross(a, b); }

inline float LengthSquared(const Vec3f& v) { return Dot(v, v); }
inline float Length(const Vec3f& v) { return std::sqrt(LengthSquared(v)); }

// v * (1 / |v|), i.e. one division then three products (reference lin_alg.h:151-156)
inline Vec3f Normalize(const Vec3f& v)
{
    const float inv_len = 1.0f / Length(v);
    return v * inv_len;
}

inline Vec3f ComponentMin(const Vec3f& a, const Vec3f& b)
{
    return Vec3f(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z));
}
inline Vec3f ComponentMax(const Vec3f& a, const Vec3f& b)
{
    return Vec3f(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z));
}

template <class T> T Clamp(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }

// 0x00RRGGBB with saturation above 1 and truncation below (reference lin_alg.h:125-132)
inline uint32 ToBGRA8(const Vec3f& c)
{
    const uchar r = c.x > 1.0f ? 255 : (uchar) (c.x * 255.0f);
    const uchar g = c.y > 1.0f ? 255 : (uchar) (c.y * 255.0f);
    const uchar b = c.z > 1.0f ? 255 : (uchar) (c.z * 255.0f);
    return uint32(r) << 16 | uint32(g) << 8 | uint32(b);
}

template <typename T> T DegToRad(const T deg) { return deg * T(0.0174532925); }
template <typename T> T RadToDeg(const T rad) { return rad * T(57.2957795131); }

// 4x4 matrix, row-vector convention: a point p transforms as p' = p * M, the translation lives
// in m_mat[3][0..2].  Constructor-style setters take their arguments in the textbook
// (column-vector) reading order and store the transpose, exactly like the reference's Set().
struct Matrix44f
{
    float m_mat[4][4];

    Matrix44f() { Identity(); }

    void Identity()
    {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                m_mat[i][j] = (i == j) ? 1.0f : 0.0f;
    }

    // rows[r][c] in textbook order -> stored transposed
    void SetRows(const float rows[4][4])
    {
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++)
                m_mat[c][r] = rows[r][c];
    }

    void Translation(float x, float y, float z)
    {
        const float rows[4][4] = { { 1, 0, 0, x }, { 0, 1, 0, y }, { 0, 0, 1, z }, { 0, 0, 0, 1 } };
        SetRows(rows);
    }

    void Scaling(float f)
    {
        const float rows[4][4] = { { f, 0, 0, 0 }, { 0, f, 0, 0 }, { 0, 0, f, 0 }, { 0, 0, 0, 1 } };
        SetRows(rows);
    }

    void RotationX(float degrees)
    {
        const float r = DegToRad(degrees), c = std::cos(r), s = std::sin(r);
        const float rows[4][4] = { { 1, 0, 0, 0 }, { 0, c, -s, 0 }, { 0, s, c, 0 }, { 0, 0, 0, 1 } };
        SetRows(rows);
    }

    void RotationY(float degrees)
    {
        const float r = DegToRad(degrees), c = std::cos(r), s = std::sin(r);
        const float rows[4][4] = { { c, 0, -s, 0 }, { 0, 1, 0, 0 }, { s, 0, c, 0 }, { 0, 0, 0, 1 } };
        SetRows(rows);
    }

    void RotationZ(float degrees)
    {
        const float r = DegToRad(degrees), c = std::cos(r), s = std::sin(r);
        const float rows[4][4] = { { c, -s, 0, 0 }, { s, c, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } };
        SetRows(rows);
    }

    // this = this * a; each element is the 4-term sum in index order 0..3
    void Multiply(const Matrix44f& a)
    {
        float old[4][4];
        std::memcpy(old, m_mat, sizeof(old));
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                m_mat[i][j] = old[i][0] * a.m_mat[0][j] + old[i][1] * a.m_mat[1][j] +
                              old[i][2] * a.m_mat[2][j] + old[i][3] * a.m_mat[3][j];
    }

    friend Matrix44f operator * (const Matrix44f& a, const Matrix44f& b)
    {
        Matrix44f r(a);
        r.Multiply(b);
        return r;
    }

    bool operator == (const Matrix44f& o) const { return std::memcmp(m_mat, o.m_mat, sizeof(m_mat)) == 0; }

    void Transf3x3(const Vec3f& p, Vec3f& out) const
    {
        const float px = p.x, py = p.y, pz = p.z;
        out.x = px * m_mat[0][0] + py * m_mat[1][0] + pz * m_mat[2][0];
        out.y = px * m_mat[0][1] + py * m_mat[1][1] + pz * m_mat[2][1];
        out.z = px * m_mat[0][2] + py * m_mat[1][2] + pz * m_mat[2][2];
    }
    void Transf3x3(Vec3f& p) const { Transf3x3(p, p); }

    void Transf4x4(const Vec3f& p, Vec3f& out) const
    {
        const float px = p.x, py = p.y, pz = p.z;
        out.x = px * m_mat[0][0] + py * m_mat[1][0] + pz * m_mat[2][0] + m_mat[3][0];
        out.y = px * m_mat[0][1] + py * m_mat[1][1] + pz * m_mat[2][1] + m_mat[3][1];
        out.z = px * m_mat[0][2] + py * m_mat[1][2] + pz * m_mat[2][2] + m_mat[3][2];
    }
    void Transf4x4(Vec3f& p) const { Transf4x4(p, p); }

    void Transpose4x4()
    {
        for (int i = 0; i < 4; i++)
            for (int j = i + 1; j < 4; j++)
                std::swap(m_mat[i][j], m_mat[j][i]);
    }

    void Transpose3x3()
    {
        for (int i = 0; i < 3; i++)
            for (int j = i + 1; j < 3; j++)
                std::swap(m_mat[i][j], m_mat[j][i]);
    }

    // Cofactor inverse.  The reference (lin_alg.h:635-687, after MESA GLU's gluInvertMatrix) writes
    // each cofactor as six signed triple products in a fixed order; fp32 addition is not
    // associative, so the same order is kept here in table form: entry k of the adjugate is
    //   sum_{t=0..5} sign[t] * m[a]*m[b]*m[c]        accumulated left to right.
    // Returns false (matrix untouched) for a singular matrix.
    bool Invert();

    // Camera matrix looking from eye to look_at (reference lin_alg.h:431-467): basis rows
    // x = up ^ z, y = z ^ x, z = normalize(eye - look_at), pre-multiplied by the translation
    // (x.eye, y.eye, z.eye)
    void BuildLookAtMatrix(const Vec3f& eye, const Vec3f& look_at, const Vec3f& up = Vec3f(0.0f, 1.0f, 0.0f))
    {
        const Vec3f zaxis = Normalize(eye - look_at);
        const Vec3f xaxis = Normalize(up ^ zaxis);
        const Vec3f yaxis = Normalize(zaxis ^ xaxis);
        Matrix44f basis;
        basis.m_mat[0][0] = xaxis.x; basis.m_mat[0][1] = xaxis.y; basis.m_mat[0][2] = xaxis.z;
        basis.m_mat[1][0] = yaxis.x; basis.m_mat[1][1] = yaxis.y; basis.m_mat[1][2] = yaxis.z;
        basis.m_mat[2][0] = zaxis.x; basis.m_mat[2][1] = zaxis.y; basis.m_mat[2][2] = zaxis.z;
        Matrix44f trans;
        trans.Translation(Dot(xaxis, eye), Dot(yaxis, eye), Dot(zaxis, eye));
        *this = trans * basis;
    }
};

inline bool Matrix44f::Invert()
{
    // One row per adjugate entry, in the order the entries are needed; each term is
    // { sign, a, b, c } meaning sign * m[a] * m[b] * m[c] with m = &m_mat[0][0]
    struct Term { signed char s; unsigned char a, b, c; };
    static const unsigned char ORDER[16] = { 0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15 };
    static const Term T[16][6] = {
        { { +1, 5, 10, 15 }, { -1, 5, 11, 14 }, { -1, 9, 6, 15 }, { +1, 9, 7, 14 }, { +1, 13, 6, 11 }, { -1, 13, 7, 10 } },
        { { -1, 4, 10, 15 }, { +1, 4, 11, 14 }, { +1, 8, 6, 15 }, { -1, 8, 7, 14 }, { -1, 12, 6, 11 }, { +1, 12, 7, 10 } },
        { { +1, 4, 9, 15 }, { -1, 4, 11, 13 }, { -1, 8, 5, 15 }, { +1, 8, 7, 13 }, { +1, 12, 5, 11 }, { -1, 12, 7, 9 } },
        { { -1, 4, 9, 14 }, { +1, 4, 10, 13 }, { +1, 8, 5, 14 }, { -1, 8, 6, 13 }, { -1, 12, 5, 10 }, { +1, 12, 6, 9 } },
        { { -1, 1, 10, 15 }, { +1, 1, 11, 14 }, { +1, 9, 2, 15 }, { -1, 9, 3, 14 }, { -1, 13, 2, 11 }, { +1, 13, 3, 10 } },
        { { +1, 0, 10, 15 }, { -1, 0, 11, 14 }, { -1, 8, 2, 15 }, { +1, 8, 3, 14 }, { +1, 12, 2, 11 }, { -1, 12, 3, 10 } },
        { { -1, 0, 9, 15 }, { +1, 0, 11, 13 }, { +1, 8, 1, 15 }, { -1, 8, 3, 13 }, { -1, 12, 1, 11 }, { +1, 12, 3, 9 } },
        { { +1, 0, 9, 14 }, { -1, 0, 10, 13 }, { -1, 8, 1, 14 }, { +1, 8, 2, 13 }, { +1, 12, 1, 10 }, { -1, 12, 2, 9 } },
        { { +1, 1, 6, 15 }, { -1, 1, 7, 14 }, { -1, 5, 2, 15 }, { +1, 5, 3, 14 }, { +1, 13, 2, 7 }, { -1, 13, 3, 6 } },
        { { -1, 0, 6, 15 }, { +1, 0, 7, 14 }, { +1, 4, 2, 15 }, { -1, 4, 3, 14 }, { -1, 12, 2, 7 }, { +1, 12, 3, 6 } },
        { { +1, 0, 5, 15 }, { -1, 0, 7, 13 }, { -1, 4, 1, 15 }, { +1, 4, 3, 13 }, { +1, 12, 1, 7 }, { -1, 12, 3, 5 } },
        { { -1, 0, 5, 14 }, { +1, 0, 6, 13 }, { +1, 4, 1, 14 }, { -1, 4, 2, 13 }, { -1, 12, 1, 6 }, { +1, 12, 2, 5 } },
        { { -1, 1, 6, 11 }, { +1, 1, 7, 10 }, { +1, 5, 2, 11 }, { -1, 5, 3, 10 }, { -1, 9, 2, 7 }, { +1, 9, 3, 6 } },
        { { +1, 0, 6, 11 }, { -1, 0, 7, 10 }, { -1, 4, 2, 11 }, { +1, 4, 3, 10 }, { +1, 8, 2, 7 }, { -1, 8, 3, 6 } },
        { { -1, 0, 5, 11 }, { +1, 0, 7, 9 }, { +1, 4, 1, 11 }, { -1, 4, 3, 9 }, { -1, 8, 1, 7 }, { +1, 8, 3, 5 } },
        { { +1, 0, 5, 10 }, { -1, 0, 6, 9 }, { -1, 4, 1, 10 }, { +1, 4, 2, 9 }, { +1, 8, 1, 6 }, { -1, 8, 2, 5 } },
    };
    const float *m = &m_mat[0][0];
    float adj[16];
    for (int k = 0; k < 16; k++)
    {
        float acc = 0.0f;
        for (int t = 0; t < 6; t++)
        {
            const Term& q = T[k][t];
            // (+-m[a]) * m[b] * m[c]: negating the first factor is exact, so this equals the
            // reference's "-m[a] * m[b] * m[c]" / "- m[a] * m[b] * m[c]" spellings bit for bit
            const float prod = m[q.a] * m[q.b] * m[q.c];
            if (t == 0)
                acc = q.s > 0 ? prod : -prod;
            else
                acc = q.s > 0 ? acc + prod : acc - prod;
        }
        adj[ORDER[k]] = acc;
    }
    float det = m[0] * adj[0] + m[1] * adj[4] + m[2] * adj[8] + m[3] * adj[12];
    if (det == 0.0f)
        return false;
    det = 1.0f / det;
    float *w = &m_mat[0][0];
    for (int i = 0; i < 16; i++)
        w[i] = adj[i] * det;
    return true;
}

#endif // RTM_HOST_LIN_ALG_H
