// TEST INFRASTRUCTURE (compiled only by oracle/build_ref.sh, only where /root/reference exists).
//
// The reference keeps its camera, primary-ray generation and pixel colouring as file-local functions of
// its front end (src/main.cpp:42-111). This translation unit includes that file unmodified — with its
// main() renamed — so that the very same functions can be called from tests and from the reference
// bench arm: they pin oracle/hagrid_oracle.c's restatement (tests/golden/frontend.npz) and they are
// the reference's half of the interactive-frame comparison.
#include <cstring>
#include <vector>

#define main hagrid_reference_main
#include "main.cpp"
#undef main

static Camera camera_from(const float* c) {
    Camera cam;
    cam.eye = vec3(c[0], c[1], c[2]); cam.right = vec3(c[3], c[4], c[5]);
    cam.up = vec3(c[6], c[7], c[8]); cam.dir = vec3(c[9], c[10], c[11]);
    return cam;
}

extern "C" {

__attribute__((visibility("default")))
void hgb_ref_gen_camera(const float* eye, const float* center, const float* up, float fov, float ratio, float* cam12) {
    const Camera cam = gen_camera(vec3(eye[0], eye[1], eye[2]), vec3(center[0], center[1], center[2]), vec3(up[0], up[1], up[2]), fov, ratio);
    const vec3 v[4] = {cam.eye, cam.right, cam.up, cam.dir};           // member order of Camera, src/main.cpp:18-23
    for (int i = 0; i < 4; i++) { cam12[3 * i] = v[i].x; cam12[3 * i + 1] = v[i].y; cam12[3 * i + 2] = v[i].z; }
}

/// gen_rays into a vector that lives across calls (the front end also reuses its vector, src/main.cpp:586)
__attribute__((visibility("default")))
const void* hgb_ref_gen_rays(const float* cam12, float clip, int w, int h) {
    static std::vector<Ray> rays;
    rays.resize(size_t(w) * h);
    gen_rays(camera_from(cam12), rays, clip, w, h);
    return rays.data();
}

/// load_model (src/main.cpp:246-275): the reference's OBJ loader + triangle setup; nullptr when it refuses the file
__attribute__((visibility("default")))
const void* hgb_ref_load_model(const char* path, int* num_tris) {
    static std::vector<Tri> tris;
    tris.clear();
    *num_tris = 0;
    if (!load_model(path, tris)) return nullptr;
    *num_tris = int(tris.size());
    return tris.data();
}

/// load_rays (src/main.cpp:277-300); nullptr when the file cannot be opened
__attribute__((visibility("default")))
const void* hgb_ref_load_rays(const char* path, float tmin, float tmax, long long* count) {
    static std::vector<Ray> rays;
    rays.clear();
    *count = 0;
    if (!load_rays(path, rays, tmin, tmax)) return nullptr;
    *count = (long long)rays.size();
    static Ray none;
    return rays.empty() ? &none : rays.data();
}

/// update_surface<mode> on a tightly packed w x h BGRA image
__attribute__((visibility("default")))
void hgb_ref_update_surface(int mode, const void* hits, float clip, int w, int h, void* bgra) {
    static std::vector<Hit> host_hits;
    host_hits.assign(static_cast<const Hit*>(hits), static_cast<const Hit*>(hits) + size_t(w) * h);
    SDL_Surface surf;
    surf.w = w; surf.h = h; surf.pitch = w * 4; surf.pixels = bgra;
    if (mode == 0)      update_surface<DisplayMode::DEPTH>(&surf, host_hits, clip, w, h);
    else if (mode == 1) update_surface<DisplayMode::GRAY_SCALE>(&surf, host_hits, clip, w, h);
    else                update_surface<DisplayMode::HEAT_MAP>(&surf, host_hits, clip, w, h);
}

}
