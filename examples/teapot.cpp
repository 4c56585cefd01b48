// examples/teapot.rs restated against include/firework.hpp (four models with vertex normals, every one rotated by
// Rotor3::from_rotation_xz(90.)).  Like suzanne.rs the crate's example dumps its scene to scenes/teapot.yml, the file the
// reference commits; `--yaml` here must parse to exactly that document (tests/test_host.py).
#include "add_obj.hpp"
#include "common.hpp"

using namespace firework;

Scene teapot_scene(const std::string& obj_file) {   // teapot.rs:66-91
    Scene scene = Scene::new_();
    auto diffuse = scene.add_material(LambertianMat::new_(ConstantTexture::new_(Vec3(0.2f, 0.8f, 0.3f))));
    const Rotor3 turn = Rotor3::from_rotation_xz(90.f);
    add_obj(scene, obj_file, diffuse, true, &turn);

    scene.set_environment(SkyEnv::default_());

    auto grey = scene.add_material(LambertianMat::with_color(Vec3::broadcast(0.5f)));
    scene.add_object(RenderObject::new_(XZRect::new_(-20.f, 20.f, -20.f, 20.f, 0.f, grey)));

    auto light = scene.add_material(EmissiveMat::with_color(Vec3::broadcast(20.f)));
    scene.add_object(RenderObject::new_(YZRect::new_(0.f, 4.f, 0.f, 4.f, -0.6f, light)).rotate(Rotor3::from_rotation_xz(-30.f)).position(0.f, 4.f, 10.f));
    return scene;
}

int main(int argc, char** argv) {   // teapot.rs:93-123
    std::string obj_file = "teapot.obj";
    for (int i = 1; i + 1 < argc; ++i)
        if (std::string(argv[i]) == "--obj") { obj_file = argv[i + 1]; for (int k = i; k + 2 < argc; ++k) argv[k] = argv[k + 2]; argc -= 2; break; }
    try {
        Scene scene = teapot_scene(obj_file);
        CameraSettings camera = CameraSettings::default_().cam_pos(Vec3(1.f, 4.f, 8.f)).look_at(Vec3(0.f, 1.f, 0.f)).field_of_view(40.f);
        Renderer renderer = Renderer::default_().width(1920).height(1080).samples(512).use_bvh(true).camera(camera);
        return run_example(argc, argv, "teapot", scene, renderer);
    } catch (const Error& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
