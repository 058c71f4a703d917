// examples/sphere.cpp -- the reference's examples/sphere.rs (lines 4-34) through the C++
// host mirror (include/rpt.hpp).  Build: make examples.  Prints the image mean and writes
// output.ppm; the reference saves output.png through the `image` crate.
#include <cstdio>

#include "../include/rpt.hpp"
using namespace rpt;

int main(int argc, char** argv) {
    Scene scene;
    scene.add(Object(sphere()));  // default red material
    scene.add(Object(plane(vec3(0.0, 1.0, 0.0), -1.0)).material(Material::diffuse(hex_color(0xAAAAAA))));
    scene.add(Light::Object(Object(sphere().scale(vec3(2.0, 2.0, 2.0)).translate(vec3(0.0, 12.0, 0.0)))
                                .material(Material::light(hex_color(0xFFFFFF), 40.0))));
    const Camera camera = Camera::look_at(vec3(-2.5, 4.0, 6.5), vec3(0.0, -0.25, 0.0), vec3(0.0, 1.0, 0.0), 0.78539816339744831);
    const uint32_t w = argc > 2 ? 96 : 960, h = argc > 2 ? 54 : 540;
    Renderer renderer(scene, camera);
    renderer.width(w).height(h).max_bounces(2).num_samples(100).seed(1);
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
