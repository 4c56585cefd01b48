// examples/part2_all.rs restated against include/firework.hpp — the BASELINE.json headline scene: 400 boxes of random height,
// an area light, glass / metal / image-textured / turbulence spheres, a medium inside glass, a cube of 1 000 small spheres and a
// thin fog over everything.  The crate's example no longer compiles against its own library (it calls a removed
// ConstantMedium::new); like firework_b200/scenes.py this uses the current Scene::add_volume (scene.rs:47-62).  The document is
// compared byte for byte with scenes/part2_all.yml in tests/test_host.py.
#include "common.hpp"
#include "scene_rng.hpp"

using namespace firework;

static Vec3 scaled(float s, Vec3 v) { return Vec3(s * v.x, s * v.y, s * v.z); }   // `f32 * Vec3`

Scene final_scene(SceneRng& rand) {   // part2_all.rs:13-80
    Scene scene = Scene::new_();

    auto ground = scene.add_material(LambertianMat::with_color(Vec3(0.48f, 0.83f, 0.53f)));
    const Vec3 origin(-10.f, 0.f, -10.f);
    for (int x = 0; x < 20; ++x) {
        for (int z = 0; z < 20; ++z) {
            const Vec3 pos = origin + Vec3((float)x, 0.f, (float)z);
            const Vec3 size(1.f, rand.rand_f32() + 0.01f, 1.f);
            scene.add_object(RenderObject::new_(Rect3d::with_size(size, ground)).position_vec(pos));
        }
    }

    auto light = scene.add_material(EmissiveMat::with_color(scaled(7.0f, Vec3::one())));
    scene.add_object(RenderObject::new_(XZRect::new_(1.23f, 4.23f, 1.47f, 4.12f, 5.54f, light)));

    auto brown = scene.add_material(LambertianMat::with_color(Vec3(0.7f, 0.3f, 0.1f)));
    scene.add_object(RenderObject::new_(Sphere::new_(0.5f, brown)).position(4.f, 4.f, 2.f));

    auto glass = scene.add_material(DielectricMat::new_(1.5f));
    scene.add_object(RenderObject::new_(Sphere::new_(0.5f, glass)).position(2.6f, 1.5f, 0.45f));
    auto metal = scene.add_material(MetalMat::new_(Vec3(0.8f, 0.8f, 0.9f), 10.0f));
    scene.add_object(RenderObject::new_(Sphere::new_(0.5f, metal)).position(0.f, 1.5f, 1.45f));

    scene.add_object(RenderObject::new_(Sphere::new_(0.7f, glass)).position(3.6f, 1.5f, 1.45f));
    scene.add_volume(RenderObject::new_(Sphere::new_(0.7f, glass)).position(3.6f, 1.5f, 1.45f), 0.2f, ConstantTexture::new_(Vec3(0.2f, 0.4f, 0.9f)));

    auto earth_mat = scene.add_material(LambertianMat::new_(ImageTexture::from_path("earthmap.jpg")));
    scene.add_object(RenderObject::new_(Sphere::new_(1.0f, earth_mat)).position(4.f, 2.f, 4.f));

    auto noise = scene.add_material(LambertianMat::new_(TurbulenceTexture::new_(5, 10.0f)));
    scene.add_object(RenderObject::new_(Sphere::new_(0.8f, noise)).position(2.2f, 2.8f, 3.f));

    auto white = scene.add_material(LambertianMat::with_color(scaled(0.73f, Vec3::one())));
    for (int i = 0; i < 1000; ++i) {
        const float a = rand.rand_f32();
        const float b = rand.rand_f32();
        const float c = rand.rand_f32();
        const Vec3 pos = scaled(1.65f, Vec3(a, b, c)) + Vec3(1.0f, 2.7f, 3.95f);
        scene.add_object(RenderObject::new_(Sphere::new_(0.1f, white)).position_vec(pos));
    }

    scene.add_volume(RenderObject::new_(Sphere::new_(5000.0f, 0)), 0.0001f, ConstantTexture::new_(Vec3::one()));
    return scene;
}

int main(int argc, char** argv) {   // part2_all.rs:82-98 (600 x 800 x 10 000 spp there; BASELINE.json quotes it at 4K x 4096)
    SceneRng rng(12345);
    Scene scene = final_scene(rng);
    CameraSettings camera = CameraSettings::default_().cam_pos(Vec3(-9.f, 3.f, -9.f)).look_at(Vec3(1.f, 3.f, 2.f)).field_of_view(25.f);
    Renderer renderer = Renderer::default_().width(600).height(800).samples(10000).use_bvh(true).camera(camera);
    return run_example(argc, argv, "part2 final", scene, renderer);
}
