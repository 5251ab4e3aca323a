// gi_loader.cpp — `.scn` scene files and OBJ meshes (sceneLoader.cpp:12-185, meshLoader.cpp:10-99), textures.
// The file formats are defined by the reference's scanf conversions, so the same conversions are used here: the
// loader reads whitespace-separated words, acts on the 14 keywords and skips anything else one
// word at a time — comments are NOT skipped as a unit, exactly like the reference.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <locale>
#include <string>
#include <vector>

#include "gi_scene.hpp"

using gi::dvec2;
using gi::dvec3;

imageTexture::imageTexture(const char* name, dvec2 t) : texture(dvec3(0, 0, 0)), fname(name), tile(t)
{
    // QImage(fname) (material.h:57): PNG and JPEG files are decoded here (gi_png.cpp, gi_jpg.cpp); anything else, and the
    // stand-in textures, come as a raw sidecar `<name>.rgba` written by scenes/stage_assets.py
    if (gi_png_decode(fname.c_str(), width, height, has_alpha, rgba)) {
        std::cout << "loading texture: " << fname << "\nsize: " << width << ", " << height << "\n";
        return;
    }
    if (gi_jpg_decode(fname.c_str(), width, height, rgba)) {
        has_alpha = false;   // QImage::hasAlphaChannel of a JPEG
        std::cout << "loading texture: " << fname << "\nsize: " << width << ", " << height << "\n";
        return;
    }
    width = height = 0; has_alpha = false; rgba.clear();
    std::string side = fname + ".rgba";
    FILE* f = std::fopen(side.c_str(), "rb");
    if (!f) { std::cout << "error while opening texture: " << fname << " (no PNG / JPEG, no sidecar " << side << ")\n"; return; }
    char magic[4];
    uint32_t hdr[3];
    if (std::fread(magic, 1, 4, f) == 4 && std::memcmp(magic, "GIRT", 4) == 0 && std::fread(hdr, 4, 3, f) == 3) {
        width = (int)hdr[0]; height = (int)hdr[1]; has_alpha = hdr[2] != 0;
        rgba.resize((size_t)width * height * 4);
        if (std::fread(rgba.data(), 1, rgba.size(), f) != rgba.size()) { width = height = 0; rgba.clear(); }
    }
    std::fclose(f);
    std::cout << "loading texture: " << fname << "\nsize: " << width << ", " << height << "\n";
}

void loadOBJ(Octree* o, const char* fname, dvec3 pos, dvec3 rotation, const Material& material)  // meshLoader.cpp:10-99
{
    struct fvec3 { float x, y, z; };
    struct fvec2 { float x, y; };
    std::vector<fvec3> verts, normals;   // the reference stores these as float (glm::vec3): positions, normals and
    std::vector<fvec2> uvs;              // uvs are narrowed to fp32 before the triangles are built
    gi::dmat3 rot = gi::eulerAngleXYZ(rotation.x, rotation.y, rotation.z);
    FILE* f = std::fopen(fname, "r");
    if (f == NULL) { std::cout << "error while opening file: " << fname << "\n"; return; }
    int faces = 0;
    std::cout << "loading mesh: " << fname << "\n";
    while (1) {
        char word[128];
        int res = std::fscanf(f, "%127s", word);
        if (res == EOF) break;
        if (std::strcmp(word, "v") == 0) {
            dvec3 v;
            std::fscanf(f, "%lf %lf %lf\n", &v.x, &v.y, &v.z);
            dvec3 w = rot * v + pos;
            verts.push_back({ (float)w.x, (float)w.y, (float)w.z });
        } else if (std::strcmp(word, "vt") == 0) {
            dvec2 uv;
            std::fscanf(f, "%lf %lf\n", &uv.x, &uv.y);
            uvs.push_back({ (float)uv.x, (float)uv.y });
        } else if (std::strcmp(word, "vn") == 0) {
            dvec3 n;
            std::fscanf(f, "%lf %lf %lf\n", &n.x, &n.y, &n.z);
            dvec3 w = rot * n;
            normals.push_back({ (float)w.x, (float)w.y, (float)w.z });
        } else if (std::strcmp(word, "f") == 0) {
            unsigned int vi[3] = { 1, 2, 3 }, ti[3], ni[3];
            int matches = std::fscanf(f, "%d%*[/]%d%*[/]%d %d%*[/]%d%*[/]%d %d%*[/]%d%*[/]%d\n", &vi[0], &ti[0], &ni[0], &vi[1], &ti[1], &ni[1], &vi[2], &ti[2], &ni[2]);
            if (matches != 9) { std::cout << "error while reading faces \n"; std::fclose(f); return; }
            vertex vx[3];
            for (int k = 0; k < 3; k++) {
                if (vi[k] - 1 >= verts.size() || ni[k] - 1 >= normals.size() || ti[k] - 1 >= uvs.size()) { std::cout << "face index out of range in " << fname << "\n"; std::fclose(f); return; }
                const fvec3& p = verts[vi[k] - 1];
                const fvec3& n = normals[ni[k] - 1];
                const fvec2& t = uvs[ti[k] - 1];
                vx[k] = vertex(dvec3(p.x, p.y, p.z), dvec3(n.x, n.y, n.z), dvec2(t.x, t.y));
            }
            o->push_back(new triangle(vx[0], vx[1], vx[2], material));
            faces++;
        }
    }
    std::cout << "read faces: " << faces << "\n";
    std::fclose(f);
}

void loadScene(Octree* o, RayTracer& r, const char* fname)  // sceneLoader.cpp:12-185
{
    std::vector<texture*> tex;
    std::vector<Material*> mats;
    std::string path = fname;
    std::size_t slash = path.find_last_of("/");
    std::string dir = path.substr(0, slash);
    FILE* f = std::fopen(fname, "r");
    if (f == NULL) { std::cout << "error while opening file: " << fname << "\n"; return; }
    std::cout << "loading scene: " << fname << "\n";
    auto mat_at = [&](int i) -> const Material* { return (i >= 0 && (size_t)i < mats.size()) ? mats[i] : nullptr; };
    while (1) {
        char word[128];
        int res = std::fscanf(f, "%127s", word);
        if (res == EOF) break;
        if (std::strcmp(word, "imTex") == 0) {
            char fn[256];
            int utile = 1, vtile = 1;
            std::fscanf(f, "%255s %d %d\n", fn, &utile, &vtile);
            tex.push_back(new imageTexture((dir + "/" + fn).c_str(), dvec2(utile, vtile)));
        } else if (std::strcmp(word, "checkerboardTex") == 0) {
            dvec3 a, b;
            int tiles = 1;
            std::fscanf(f, "%lf %lf %lf %lf %lf %lf %d\n", &a.x, &a.y, &a.z, &b.x, &b.y, &b.z, &tiles);
            tex.push_back(new checkerboard(tiles, a, b));
        } else if (std::strcmp(word, "colorTex") == 0) {
            dvec3 col;
            std::fscanf(f, "%lf %lf %lf\n", &col.x, &col.y, &col.z);
            tex.push_back(new texture(col));
        } else if (std::strcmp(word, "mat") == 0) {
            int dif = 0, em = 0;
            double rough = 1, op = 1, ior = 1;  // a 4-field line leaves IOR at 1 here (uninitialised in the reference, §A.10)
            std::fscanf(f, "%d %d %lf %lf %lf\n", &dif, &em, &rough, &op, &ior);
            if (dif < 0 || em < 0 || (size_t)dif >= tex.size() || (size_t)em >= tex.size()) { std::cout << "mat: texture index out of range\n"; continue; }
            mats.push_back(new Material(tex[dif], tex[em], rough, op, ior));
        } else if (std::strcmp(word, "multiMat") == 0) {
            char s[128];
            int length = 0;
            std::fscanf(f, "%127[0123456789 ]%n\n", s, &length);  // parsed and ignored, like the reference
        } else if (std::strcmp(word, "mesh") == 0) {
            char fn[256];
            dvec3 pos, rot;
            int mat = 0;
            std::fscanf(f, "%255s %lf %lf %lf %lf %lf %lf %d\n", fn, &pos.x, &pos.y, &pos.z, &rot.x, &rot.y, &rot.z, &mat);
            if (!mat_at(mat)) { std::cout << "mesh: material index out of range\n"; continue; }
            loadOBJ(o, (dir + "/" + fn).c_str(), pos, rot, *mat_at(mat));
        } else if (std::strcmp(word, "sphere") == 0) {
            dvec3 pos;
            double rad = 1;
            int mat = 0;
            std::fscanf(f, "%lf %lf %lf %lf %d\n", &pos.x, &pos.y, &pos.z, &rad, &mat);
            if (!mat_at(mat)) { std::cout << "sphere: material index out of range\n"; continue; }
            o->push_back(new sphere(pos, rad, *mat_at(mat)));
        } else if (std::strcmp(word, "box") == 0) {
            dvec3 pos, size, rot;
            int mat = 0;
            std::fscanf(f, "%lf %lf %lf %lf %lf %lf %lf %lf %lf %d\n", &pos.x, &pos.y, &pos.z, &size.x, &size.y, &size.z, &rot.x, &rot.y, &rot.z, &mat);
            if (!mat_at(mat)) { std::cout << "box: material index out of range\n"; continue; }
            boxMesh(o, pos, size, rot, *mat_at(mat));
        } else if (std::strcmp(word, "light") == 0) {
            dvec3 pos, col;
            double rad = 0;
            std::fscanf(f, "%lf %lf %lf %lf %lf %lf %lf\n", &pos.x, &pos.y, &pos.z, &col.x, &col.y, &col.z, &rad);
            o->push_back(new Light(pos, dvec3(0, 0, 0), col, rad));
        } else if (std::strcmp(word, "heightFog") == 0) {
            dvec3 pos, size, col;   // sceneLoader.cpp:150-159
            double density = 0, scatter = 0, scale = 0;
            std::fscanf(f, "%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf\n", &pos.x, &pos.y, &pos.z, &size.x, &size.y, &size.z, &col.x, &col.y, &col.z, &density, &scatter, &scale);
            o->push_back(new HeightFog(pos, size, col, density, scatter, (int)scale));
        } else if (std::strcmp(word, "photons") == 0) {
            std::fscanf(f, "%d %d\n", &r.photons, &r.photon_depth);
        } else if (std::strcmp(word, "samples") == 0) {
            std::fscanf(f, "%d %d %lf\n", &r.min_samples, &r.max_samples, &r.noise_thresh);
        } else if (std::strcmp(word, "ambient") == 0) {
            std::fscanf(f, "%lf %lf %lf\n", &r.ambient.x, &r.ambient.y, &r.ambient.z);
        } else if (std::strcmp(word, "camera") == 0) {
            dvec3 lookAt;
            std::fscanf(f, "%lf %lf %lf %lf %lf %lf\n", &r._camera.pos.x, &r._camera.pos.y, &r._camera.pos.z, &lookAt.x, &lookAt.y, &lookAt.z);
            r._camera.setDir(lookAt - r._camera.pos);
        }
    }
    std::cout << "finished scene loading\n";
    std::fclose(f);
}
