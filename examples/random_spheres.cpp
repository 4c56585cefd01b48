// examples/random_spheres.rs restated against include/firework.hpp.  The crate's example draws its scene from
// tiny_rng::Rng::new(12345), a crate that is not vendored with the reference; like firework_b200/scenes.py this uses a
// SplitMix64 stream with the same seed in its place, so both mirrors generate the same 485 spheres (the document is compared
// byte for byte with the Python mirror's in tests/test_host.py).
#include "common.hpp"
#include "scene_rng.hpp"

using namespace firework;

Scene random_scene(SceneRng& rand) {   // random_spheres.rs:14-67
    Scene scene = Scene::new_();
    auto checker_mat = scene.add_material(LambertianMat::new_(CheckerTexture::with_colors(Vec3(0.2f, 0.4f, 0.1f), Vec3(0.9f, 0.9f, 0.9f), 10.0f)));
    scene.add_object(RenderObject::new_(Sphere::new_(1000.0f, checker_mat)).position(0.f, -1000.f, -1.f));

    for (int x = -11; x < 11; ++x) {
        for (int y = -11; y < 11; ++y) {
            const float cx = (float)x + 0.9f * rand.rand_f32();
            const float cz = (float)y + 0.9f * rand.rand_f32();
            const Vec3 center(cx, 0.2f, cz);
            if ((center - Vec3(4.0f, 0.2f, 0.9f)).mag() > 0.9f) {
                const float pick = rand.rand_f32();
                MaterialIdx mat;
                if (pick < 0.8f) {
                    const float r = rand.rand_f32() * rand.rand_f32();
                    const float g = rand.rand_f32() * rand.rand_f32();
                    const float b = rand.rand_f32() * rand.rand_f32();
                    mat = scene.add_material(LambertianMat::with_color(Vec3(r, g, b)));
                } else if (pick < 0.95f) {
                    const float r = 0.5f * (1.0f + rand.rand_f32());
                    const float g = 0.5f * (1.0f + rand.rand_f32());
                    const float b = 0.5f * (1.0f + rand.rand_f32());
                    const float rough = 0.5f * rand.rand_f32();
                    mat = scene.add_material(MetalMat::new_(Vec3(r, g, b), rough));
                } else {
                    mat = scene.add_material(DielectricMat::new_(1.5f));
                }
                scene.add_object(RenderObject::new_(Sphere::new_(0.2f, mat)).position_vec(center));
            }
        }
    }

    auto glass = scene.add_material(DielectricMat::new_(1.5f));
    auto diffuse = scene.add_material(LambertianMat::with_color(Vec3(0.4f, 0.2f, 0.1f)));
    auto metal = scene.add_material(MetalMat::new_(Vec3(0.7f, 0.6f, 0.5f), 0.0f));
    scene.add_object(RenderObject::new_(Sphere::new_(1.0f, glass)).position(0.f, 1.f, 0.f));
    scene.add_object(RenderObject::new_(Sphere::new_(1.0f, diffuse)).position(-4.f, 1.f, 0.f));
    scene.add_object(RenderObject::new_(Sphere::new_(1.0f, metal)).position(4.f, 1.f, 0.f));

    scene.set_environment(SkyEnv::default_());
    return scene;
}

int main(int argc, char** argv) {   // random_spheres.rs:69-100
    SceneRng rng(12345);
    Scene scene = random_scene(rng);
    CameraSettings camera = CameraSettings::default_().cam_pos(Vec3(13.f, 2.f, 3.f)).look_at(Vec3::zero()).aperture(0.1f);
    Renderer renderer = Renderer::default_().width(960).height(540).samples(32).use_bvh(true).camera(camera);
    return run_example(argc, argv, "Random Spheres", scene, renderer);
}
