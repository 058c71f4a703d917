// examples/fractal_spheres.cpp -- the reference's examples/fractal_spheres.rs (lines 3-77) through the C++
// host mirror (include/rpt.hpp): five kd-trees over whole spheres (1, 6, 30, 150, 750 of them), one per
// recursion level.  Build: make examples.  Writes output.ppm (the reference saves output.png).
#include <cmath>
#include <cstdio>

#include "../include/rpt.hpp"
using namespace rpt;

static void gen(std::vector<std::vector<Shape>>& spheres, Vec3 p, double rad, size_t depth, int last_dir) {
    spheres[depth].push_back(sphere().scale(vec3(rad, rad, rad)).translate(p));
    if (depth == spheres.size() - 1) return;
    const double disp = rad * 7.0 / 5.0;
    const double dx[6] = {disp, -disp, 0.0, 0.0, 0.0, 0.0};
    const double dy[6] = {0.0, 0.0, disp, -disp, 0.0, 0.0};
    const double dz[6] = {0.0, 0.0, 0.0, 0.0, disp, -disp};
    for (int i = 0; i < 6; i++)
        if (last_dir < 0 || i != (last_dir ^ 1))
            gen(spheres, vec3(p.x + dx[i], p.y + dy[i], p.z + dz[i]), rad * 2.0 / 5.0, depth + 1, i);
}

int main(int argc, char** argv) {
    const uint32_t colors[5] = {0x264653, 0x2A9D8F, 0xE9C46A, 0xF4A261, 0xE76F51};
    std::vector<std::vector<Shape>> spheres(5);
    gen(spheres, vec3(0.0, 0.0, 0.0), 1.0, 0, -1);

    Scene scene;
    for (size_t i = 0; i < spheres.size(); i++) {
        std::printf("Level %zu: %zu spheres\n", i, spheres[i].size());
        scene.add(Object(KdTree(spheres[i])).material(Material::specular(hex_color(colors[i]), 0.25)));
    }
    scene.add(Object(plane(vec3(0.0, 0.0, 1.0), -6.0)).material(Material::diffuse(hex_color(0xffcccc))));
    scene.add(Light::Ambient(vec3(0.02, 0.02, 0.02)));
    scene.add(Light::Directional(vec3(0.6, 0.6, 0.6), normalize(vec3(0.0, -0.65, -1.0))));
    scene.add(Light::Point(vec3(100.0, 100.0, 100.0), vec3(0.0, 5.0, 5.0)));

    Camera camera;
    camera.eye = vec3(2.0, 3.5, 7.0);
    camera.direction = normalize(vec3(-0.285714, -0.5, -1.0));
    camera.up = normalize(vec3(0.0, 1.0, -0.5));
    camera.fov = 0.52359877559829887;  // FRAC_PI_6
    const uint32_t w = argc > 2 ? 96 : 800, h = argc > 2 ? 72 : 600;
    Renderer renderer(scene, camera);
    renderer.width(w).height(h).seed(1);
    const std::vector<uint8_t> rgb = renderer.render();
    double mean = 0;
    for (uint8_t v : rgb) mean += v;
    std::printf("rendered %ux%u, %llu segments in %.2f ms on the GPU, mean byte %.3f\n", w, h,
                (unsigned long long)renderer.stats.segments, renderer.stats.gpu_ms, mean / rgb.size());
    if (FILE* f = std::fopen(argc > 1 ? argv[1] : "output.ppm", "wb")) {
        std::fprintf(f, "P6\n%u %u\n255\n", w, h);
        std::fwrite(rgb.data(), 1, rgb.size(), f);
        std::fclose(f);
    }
    return 0;
}
