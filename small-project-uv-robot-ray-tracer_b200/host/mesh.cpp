// mesh.cpp -- binary glTF (.glb) room loader and floor-height estimate.
// Behaviour follows /root/reference/mesh.cpp:5-136: only meshes[0].primitives[0] is read,
// POSITION (float VEC3) is de-indexed through u16 or u32 indices into 64-byte Tri records,
// then the floor height is estimated and the BVH built.  Opt-in (Mesh::loadWholeScene):
// every triangle primitive of the default scene's node tree, node transforms applied.  The container is parsed directly
// (12-byte header, JSON chunk, BIN chunk) instead of through tinygltf.
#include "precomp.h"
#include "json_min.h"
#include <fstream>

namespace Tmpl8 {

static std::string g_assetRoot;
static bool g_assetRootInit = false;

const std::string& AssetRoot()
{
    if (!g_assetRootInit) {
        const char* e = getenv("UVRT_ASSET_ROOT");
        if (e && *e) {
            g_assetRoot = e;
            if (g_assetRoot.back() != '/') g_assetRoot += '/';
        }
        g_assetRootInit = true;
    }
    return g_assetRoot;
}

void SetAssetRoot(const std::string& dir)
{
    g_assetRoot = dir;
    if (!g_assetRoot.empty() && g_assetRoot.back() != '/') g_assetRoot += '/';
    g_assetRootInit = true;
}

Mesh::~Mesh() { Release(); }

void Mesh::Release()
{
    delete bvh; bvh = 0;
    free(triangles); triangles = 0;
    delete[] vertices; vertices = 0;
    delete[] uvcoords; uvcoords = 0;
    triangleCount = 0;
    vertexCount = 0;
    loadedMesh = false;
}

namespace {

uint32_t rd32(const std::vector<char>& b, size_t off)
{
    uint32_t v;
    memcpy(&v, &b[off], 4);
    return v;
}

struct AccessorView {
    const unsigned char* data = nullptr;
    size_t count = 0, stride = 0;
    int componentType = 0;
};

// resolves accessor -> bufferView -> the GLB's BIN chunk
bool resolve(const uvrt_json::Value& js, const unsigned char* bin, size_t binLen, long long accessorIdx,
             size_t elemBytes, AccessorView& out, std::string& err)
{
    const uvrt_json::Value& acc = js["accessors"][(size_t)accessorIdx];
    if (accessorIdx < 0 || acc.kind != uvrt_json::Value::Object) { err = "accessor missing"; return false; }
    const uvrt_json::Value& bv = js["bufferViews"][(size_t)acc["bufferView"].as_int(-1)];
    if (bv.kind != uvrt_json::Value::Object) { err = "bufferView missing"; return false; }
    if (bv["buffer"].as_int(0) != 0) { err = "only the embedded GLB buffer is supported"; return false; }
    // every number below comes from the file: negative, fractional-huge or wrapping values must not get through
    const long long o1 = bv["byteOffset"].as_int(0), o2 = acc["byteOffset"].as_int(0);
    const long long cnt = acc["count"].as_int(0), strideIn = bv["byteStride"].as_int(0);
    if (o1 < 0 || o2 < 0 || cnt < 0 || strideIn < 0 || (unsigned long long)o1 > binLen || (unsigned long long)o2 > binLen ||
        (unsigned long long)cnt > binLen || (unsigned long long)strideIn > binLen) { err = "accessor exceeds the BIN chunk"; return false; }
    size_t off = (size_t)o1 + (size_t)o2;
    out.count = (size_t)cnt;
    out.componentType = (int)acc["componentType"].as_int(0);
    out.stride = (size_t)strideIn;
    if (out.stride == 0) out.stride = elemBytes;
    if (out.count && (off > binLen || (out.count - 1) > (binLen - off) / out.stride || off + (out.count - 1) * out.stride + elemBytes > binLen)) {
        err = "accessor exceeds the BIN chunk";
        return false;
    }
    out.data = bin + off;
    return true;
}

} // namespace

void Mesh::LoadMesh()
{
    try {
        LoadMeshImpl();
    } catch (const std::exception& e) {
        // out-of-memory or a length the file lied about: a load failure like any other (the reference would crash)
        Release();
        lastError = std::string("rooms/") + modelFile + ".glb: " + e.what();
        printf("Failed to parse glTF: %s\n", lastError.c_str());
    }
}

void Mesh::LoadMeshImpl()
{
    std::cout << "Loading mesh " << std::endl;
    Release();
    lastError.clear();
    std::string path = AssetRoot() + "rooms/" + modelFile + ".glb";
    std::ifstream f(path, std::ios::binary);
    if (!f) { lastError = "cannot open " + path; printf("Failed to parse glTF: %s\n", lastError.c_str()); return; }
    std::vector<char> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto bail = [&](const std::string& why) { lastError = path + ": " + why; printf("Failed to parse glTF: %s\n", lastError.c_str()); };
    if (file.size() < 20 || memcmp(file.data(), "glTF", 4) != 0) return bail("not a binary glTF file");
    if (rd32(file, 4) != 2) return bail("unsupported glTF container version");
    size_t total = std::min<size_t>(rd32(file, 8), file.size());
    size_t jsonLen = rd32(file, 12);
    if (rd32(file, 16) != 0x4E4F534Au || 20 + jsonLen > total) return bail("missing JSON chunk");
    uvrt_json::Value js;
    std::string jerr;
    if (!uvrt_json::Parser(&file[20], jsonLen).parse(js, jerr)) return bail("JSON: " + jerr);
    size_t binHdr = 20 + jsonLen;
    if (binHdr + 8 > total || rd32(file, binHdr + 4) != 0x004E4942u) return bail("missing BIN chunk");
    size_t binLen = rd32(file, binHdr);
    if (binHdr + 8 + binLen > total) return bail("truncated BIN chunk");
    const unsigned char* bin = (const unsigned char*)&file[binHdr + 8];

    // de-indexed positions (9 floats per triangle) and uvs (6 per triangle) of everything that is loaded
    std::vector<float> outPos, outUv;
    std::string perr;
    // one primitive; `xf` (column-major 4x4, double) is applied to the positions when non-null
    auto add_primitive = [&](const uvrt_json::Value& prim, const double* xf, bool strict) -> bool {
        if (prim.kind != uvrt_json::Value::Object) { perr = "primitive missing"; return false; }
        const long long mode = prim["mode"].is_number() ? prim["mode"].as_int() : 4;
        if (!strict && mode != 4) return true;                       // points, lines, strips: not surfaces
        if (!prim["attributes"]["POSITION"].is_number()) { perr = "primitive needs POSITION"; return false; }
        if (strict && !prim["indices"].is_number()) { perr = "primitive needs POSITION and indices"; return false; }
        AccessorView pos, idx, uv;
        std::string err;
        if (!resolve(js, bin, binLen, prim["attributes"]["POSITION"].as_int(), 12, pos, err)) { perr = "POSITION: " + err; return false; }
        if (pos.componentType != 5126) { perr = "POSITION must be float"; return false; }
        const bool indexed = prim["indices"].is_number();
        size_t idxBytes = 0;
        if (indexed) {
            long long idxAcc = prim["indices"].as_int();
            int idxType = (int)js["accessors"][(size_t)idxAcc]["componentType"].as_int(0);
            // the reference reads u16 and u32 (mesh.cpp:44-51); u8 is the third type glTF allows
            if (idxType == 5123) idxBytes = 2;
            else if (idxType == 5125) idxBytes = 4;
            else if (idxType == 5121 && !strict) idxBytes = 1;
            else { perr = "indices must be u16 or u32"; return false; }
            if (!resolve(js, bin, binLen, idxAcc, idxBytes, idx, err)) { perr = "indices: " + err; return false; }
        }
        bool haveUv = prim["attributes"]["TEXCOORD_0"].is_number() &&
                      resolve(js, bin, binLen, prim["attributes"]["TEXCOORD_0"].as_int(), 8, uv, err) && uv.componentType == 5126;
        const size_t nCorners = indexed ? idx.count : pos.count;
        const size_t nTri = nCorners / 3;
        auto index_at = [&](size_t k) -> size_t {
            if (!indexed) return k;
            const unsigned char* p = idx.data + k * idx.stride;
            if (idxBytes == 1) return *p;
            if (idxBytes == 2) { uint16_t v; memcpy(&v, p, 2); return v; }
            uint32_t v; memcpy(&v, p, 4); return v;
        };
        const size_t base = outPos.size() / 9;
        outPos.resize((base + nTri) * 9);
        outUv.resize((base + nTri) * 6, 0.0f);
        for (size_t t = 0; t < nTri; t++) {
            for (int c = 0; c < 3; c++) {
                size_t v = index_at(t * 3 + c);
                if (v >= pos.count) { perr = "vertex index out of range"; return false; }
                float p[3];
                memcpy(p, pos.data + v * pos.stride, 12);
                if (xf) {
                    const double x = p[0], y = p[1], z = p[2];
                    p[0] = (float)(xf[0] * x + xf[4] * y + xf[8] * z + xf[12]);
                    p[1] = (float)(xf[1] * x + xf[5] * y + xf[9] * z + xf[13]);
                    p[2] = (float)(xf[2] * x + xf[6] * y + xf[10] * z + xf[14]);
                }
                memcpy(&outPos[(base + t) * 9 + c * 3], p, 12);
                if (haveUv && v < uv.count) memcpy(&outUv[(base + t) * 6 + c * 2], uv.data + v * uv.stride, 8);
            }
        }
        return true;
    };

    if (!loadWholeScene) {
        // the reference: meshes[0].primitives[0] only, node transforms ignored (mesh.cpp:28)
        const uvrt_json::Value& prim = js["meshes"][0]["primitives"][0];
        if (prim.kind != uvrt_json::Value::Object) return bail("meshes[0].primitives[0] missing");
        if (!add_primitive(prim, nullptr, true)) return bail(perr);
    } else {
        // every triangle primitive of every mesh instanced by the default scene's node tree, with the
        // nodes' transforms (matrix or translation * rotation * scale) applied
        struct Frame { size_t node; double m[16]; int depth; };
        auto mul = [](const double* a, const double* b, double* o) {
            for (int c = 0; c < 4; c++)
                for (int r = 0; r < 4; r++) {
                    double v = 0;
                    for (int k = 0; k < 4; k++) v += a[k * 4 + r] * b[c * 4 + k];
                    o[c * 4 + r] = v;
                }
        };
        auto local = [](const uvrt_json::Value& n, double* m) {
            for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0 : 0.0;
            if (n["matrix"].size() == 16) {
                for (int i = 0; i < 16; i++) m[i] = n["matrix"][(size_t)i].num;
                return;
            }
            double t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, sc[3] = {1, 1, 1};
            if (n["translation"].size() == 3) for (int i = 0; i < 3; i++) t[i] = n["translation"][(size_t)i].num;
            if (n["rotation"].size() == 4) for (int i = 0; i < 4; i++) q[i] = n["rotation"][(size_t)i].num;
            if (n["scale"].size() == 3) for (int i = 0; i < 3; i++) sc[i] = n["scale"][(size_t)i].num;
            const double x = q[0], y = q[1], z = q[2], w = q[3];
            const double r[9] = {1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w),      // column 0
                                 2 * (x * y - z * w), 1 - 2 * (x * x + z * z), 2 * (y * z + x * w),      // column 1
                                 2 * (x * z + y * w), 2 * (y * z - x * w), 1 - 2 * (x * x + y * y)};     // column 2
            for (int c = 0; c < 3; c++)
                for (int k = 0; k < 3; k++) m[c * 4 + k] = r[c * 3 + k] * sc[c];
            m[12] = t[0]; m[13] = t[1]; m[14] = t[2];
        };
        const uvrt_json::Value& nodes = js["nodes"];
        std::vector<Frame> stack;
        const uvrt_json::Value& roots = js["scenes"][(size_t)js["scene"].as_int(0)]["nodes"];
        double ident[16];
        for (int i = 0; i < 16; i++) ident[i] = (i % 5 == 0) ? 1.0 : 0.0;
        if (roots.size() == 0) {
            // no scene graph: every mesh once, untransformed
            for (size_t mi = 0; mi < js["meshes"].size(); mi++)
                for (size_t pi = 0; pi < js["meshes"][mi]["primitives"].size(); pi++)
                    if (!add_primitive(js["meshes"][mi]["primitives"][pi], nullptr, false)) return bail(perr);
        }
        for (size_t i = roots.size(); i-- > 0;) {
            Frame f;
            f.node = (size_t)roots[i].as_int(-1);
            memcpy(f.m, ident, sizeof ident);
            f.depth = 0;
            stack.push_back(f);
        }
        size_t visited = 0;
        while (!stack.empty()) {
            Frame f = stack.back();
            stack.pop_back();
            const uvrt_json::Value& n = nodes[f.node];
            if (n.kind != uvrt_json::Value::Object) return bail("scene references a missing node");
            if (f.depth > 256 || ++visited > 16 * (nodes.size() + 1)) return bail("node hierarchy is cyclic");
            double loc[16], world[16];
            local(n, loc);
            mul(f.m, loc, world);
            if (n["mesh"].is_number()) {
                const uvrt_json::Value& prims = js["meshes"][(size_t)n["mesh"].as_int()]["primitives"];
                for (size_t pi = 0; pi < prims.size(); pi++)
                    if (!add_primitive(prims[pi], world, false)) return bail(perr);
            }
            const uvrt_json::Value& kids = n["children"];
            for (size_t i = kids.size(); i-- > 0;) {
                Frame c;
                c.node = (size_t)kids[i].as_int(-1);
                memcpy(c.m, world, sizeof world);
                c.depth = f.depth + 1;
                stack.push_back(c);
            }
        }
    }

    const size_t nTri = outPos.size() / 9;
    if (nTri == 0) return bail("no triangles");
    // NaN / infinite coordinates would index out of the builder's bins (the reference does not check either)
    for (float v : outPos)
        if (!std::isfinite(v)) return bail("a vertex coordinate is not finite");
    if (nTri > 0x3fffffffu) return bail("too many triangles");
    void* mem = nullptr;
    if (posix_memalign(&mem, 64, sizeof(Tri) * nTri) != 0) return bail("out of memory");
    memset(mem, 0, sizeof(Tri) * nTri);   // the reference leaves the pad lanes uninitialised
    triangles = (Tri*)mem;
    vertices = new float[nTri * 9];
    uvcoords = new float[nTri * 6]();
    memcpy(vertices, outPos.data(), sizeof(float) * nTri * 9);
    memcpy(uvcoords, outUv.data(), sizeof(float) * nTri * 6);
    for (size_t t = 0; t < nTri; t++) {
        const float* p = &vertices[t * 9];
        triangles[t].vertex0 = make_float3_strict(p[0], p[1], p[2]);
        triangles[t].vertex1 = make_float3_strict(p[3], p[4], p[5]);
        triangles[t].vertex2 = make_float3_strict(p[6], p[7], p[8]);
    }
    vertexCount = (int)(nTri * 9);
    triangleCount = (int)nTri;

    DetermineFloorHeight();

    std::cout << "Vertex count: " << vertexCount << " triangle count: " << triangleCount << std::endl;
    if (buildBvhOnLoad) {
        bvh = new BVH(this);
        std::cout << "BVH size: " << bvh->nodesUsed << std::endl;
    }
    BindMesh();
    loadedMesh = true;
}

void Mesh::SetTriangles(const Tri* tris, int n, bool buildBvh)
{
    Release();
    lastError.clear();
    if (n <= 0) return;
    for (int t = 0; t < n; t++) {
        const float* v = reinterpret_cast<const float*>(&tris[t]);
        for (int k = 0; k < 12; k++)
            if ((k & 3) != 3 && !std::isfinite(v[k])) { lastError = "SetTriangles: a vertex coordinate is not finite"; return; }
    }
    void* mem = nullptr;
    if (posix_memalign(&mem, 64, sizeof(Tri) * (size_t)n) != 0) return;
    memcpy(mem, tris, sizeof(Tri) * (size_t)n);
    triangles = (Tri*)mem;
    triangleCount = n;
    vertexCount = n * 9;
    vertices = new float[(size_t)n * 9];
    uvcoords = new float[(size_t)n * 6]();
    for (int t = 0; t < n; t++) {
        const float3_strict* src[3] = {&triangles[t].vertex0, &triangles[t].vertex1, &triangles[t].vertex2};
        for (int c = 0; c < 3; c++) {
            vertices[t * 9 + c * 3 + 0] = src[c]->x;
            vertices[t * 9 + c * 3 + 1] = src[c]->y;
            vertices[t * 9 + c * 3 + 2] = src[c]->z;
        }
    }
    DetermineFloorHeight();
    if (buildBvh && buildBvhOnLoad) bvh = new BVH(this);
    loadedMesh = true;
}

// Floor = centre of the fullest of 48 height bins between the lowest vertex (at most 0) and 0.
// Same strict comparisons and fp32 bin edges as mesh.cpp:100-136 of the reference (whose
// maxVal is never raised above 0: rooms entirely above y = 0 get floor 0).
void Mesh::DetermineFloorHeight()
{
    const int binCount = 48;
    const int nVerts = vertexCount / 3;
    float maxVal = 0.0f, minVal = 0.0f;
    for (int i = 0; i < nVerts; i++) {
        float y = vertices[i * 3 + 1];
        if (y < minVal) minVal = y;
    }
    const float range = maxVal - minVal;
    float lo[binCount], hi[binCount];
    for (int j = 0; j < binCount; j++) {
        lo[j] = j * range / binCount + minVal;
        hi[j] = (j + 1) * range / binCount + minVal;
    }
    int hist[binCount] = {0};
    for (int i = 0; i < nVerts; i++) {
        const float y = vertices[i * 3 + 1];
        for (int j = 0; j < binCount; j++)
            if (lo[j] < y && y < hi[j]) hist[j]++;
    }
    int maxCount = 0, maxIndex = -1;
    for (int j = 0; j < binCount; j++)
        if (hist[j] > maxCount) { maxIndex = j; maxCount = hist[j]; }
    floorHeight = (maxIndex + 0.5f) * range / binCount + minVal;
}

void Mesh::BindMesh()
{
    // The reference creates the GL vertex / uv / colour buffers here (mesh.cpp:138-200).
    // The colour buffer lives in libuvrt (UVRT_BUF_COLOR); there is nothing to bind.
}

} // namespace Tmpl8
