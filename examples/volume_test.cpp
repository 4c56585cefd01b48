// examples/volume_test.rs restated against include/firework.hpp (Scene::add_volume -> ConstantMedium + IsotropicMat, glass, metal).
#include "common.hpp"

using namespace firework;

Scene volume_scene() {   // volume_test.rs:11-44
    Scene scene = Scene::new_();

    auto glass = scene.add_material(DielectricMat::new_(1.5f));
    auto diffuse = scene.add_material(LambertianMat::with_color(Vec3(0.8f, 0.8f, 0.8f)));
    auto metal = scene.add_material(MetalMat::new_(Vec3(0.7f, 0.7f, 0.7f), 0.0f));
    (void)metal;

    scene.add_volume(RenderObject::new_(Sphere::new_(1.0f, diffuse)).position(0.f, 1.f, 0.f), 0.5f, ConstantTexture::from_rgb(0.5f, 0.0f, 0.8f));
    scene.add_object(RenderObject::new_(Sphere::new_(1.01f, glass)).position(0.f, 1.f, 1.f));
    scene.add_object(RenderObject::new_(XZRect::new_(-100.f, 100.f, -100.f, 100.f, 0.f, diffuse)));

    auto light = scene.add_material(EmissiveMat::with_color(Vec3::broadcast(8.f)));
    scene.add_object(RenderObject::new_(YZRect::new_(0.f, 20.f, 0.f, 10.f, -3.f, light)).rotate(Rotor3::from_rotation_xz(-30.f)).position(0.f, 0.f, -10.f));

    scene.set_environment(SkyEnv::default_());
    return scene;
}

int main(int argc, char** argv) {   // volume_test.rs:46-80
    Scene scene = volume_scene();
    CameraSettings camera = CameraSettings::default_().cam_pos(Vec3(0.f, 2.f, -10.f)).look_at(Vec3::zero());
    Renderer renderer = Renderer::default_().width(960).height(540).samples(2048).camera(camera);
    return run_example(argc, argv, "volume", scene, renderer);
}
