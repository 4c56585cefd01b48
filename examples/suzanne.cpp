// examples/suzanne.rs restated against include/firework.hpp.  The crate's example writes `serde_yaml::to_string(&scene)` to
// scenes/suzanne.yml before it renders — the file the reference commits — so `--yaml` here must parse to exactly that document
// (tests/test_host.py: same objects, same f32 values; the two serialisers only differ in how many digits they print).
#include "add_obj.hpp"
#include "common.hpp"

using namespace firework;

Scene suzanne_scene(const std::string& obj_file) {   // suzanne.rs:53-79
    Scene scene = Scene::new_();
    auto diffuse = scene.add_material(LambertianMat::new_(ConstantTexture::new_(Vec3(0.8f, 0.2f, 0.3f))));
    add_obj(scene, obj_file, diffuse, false, nullptr);

    scene.set_environment(SkyEnv::default_());

    auto blue = scene.add_material(LambertianMat::with_color(Vec3(0.2f, 0.2f, 0.8f)));
    scene.add_object(RenderObject::new_(XZRect::new_(-100.f, 100.f, -100.f, 100.f, 0.f, blue)).position(0.f, -1.f, 0.f));

    auto light = scene.add_material(EmissiveMat::with_color(Vec3::broadcast(8.f)));
    scene.add_object(RenderObject::new_(YZRect::new_(0.f, 20.f, 0.f, 20.f, -3.f, light)).rotate(Rotor3::from_rotation_xz(-30.f)).position(0.f, 4.f, 10.f));
    return scene;
}

int main(int argc, char** argv) {   // suzanne.rs:81-110
    std::string obj_file = "suzanne.obj";
    for (int i = 1; i + 1 < argc; ++i)
        if (std::string(argv[i]) == "--obj") { obj_file = argv[i + 1]; for (int k = i; k + 2 < argc; ++k) argv[k] = argv[k + 2]; argc -= 2; break; }
    try {
        Scene scene = suzanne_scene(obj_file);
        CameraSettings camera = CameraSettings::default_().cam_pos(Vec3(1.f, 2.5f, 5.f)).look_at(Vec3(0.f, 0.f, 0.f)).field_of_view(40.f);
        Renderer renderer = Renderer::default_().width(960).height(540).samples(512).use_bvh(true).camera(camera);
        return run_example(argc, argv, "suzanne", scene, renderer);
    } catch (const Error& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
