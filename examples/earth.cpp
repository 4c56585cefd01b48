// examples/earth.rs restated against include/firework.hpp (image textures, a rotated light, SkyEnv).  The crate's example
// decodes the files itself (`image::open`) and hands the pixels to ImageTexture::new; here the texture carries its path — what
// the serde document holds (texture.rs:251-278) — and the library decodes it at render time (fw_image_load).
#include "common.hpp"

using namespace firework;

Scene earth_scene() {   // earth.rs:12-40
    Scene scene = Scene::new_();
    auto earth_mat = scene.add_material(LambertianMat::new_(ImageTexture::from_path("earthmap.jpg")));
    auto uv_image_mat = scene.add_material(LambertianMat::new_(ImageTexture::from_path("uvmap.png")));

    scene.add_object(RenderObject::new_(Sphere::new_(0.25f, earth_mat)));
    scene.add_object(RenderObject::new_(Sphere::new_(0.25f, uv_image_mat)).position(1.f, 0.f, 0.f));
    scene.add_object(RenderObject::new_(Sphere::new_(0.25f, earth_mat)).position(0.f, 1.f, 0.f));
    scene.add_object(RenderObject::new_(Sphere::new_(0.25f, earth_mat)).position(0.f, 0.f, 1.f));

    auto grey = scene.add_material(LambertianMat::with_color(Vec3::broadcast(0.5f)));
    scene.add_object(RenderObject::new_(XZRect::new_(-100.f, 100.f, -100.f, 100.f, 0.f, grey)));

    auto light = scene.add_material(EmissiveMat::with_color(Vec3::broadcast(8.f)));
    scene.add_object(RenderObject::new_(YZRect::new_(0.f, 20.f, 0.f, 10.f, -3.f, light)).rotate(Rotor3::from_rotation_xz(-30.f)).position(0.f, 0.f, -10.f));

    scene.set_environment(SkyEnv::default_());
    return scene;
}

int main(int argc, char** argv) {   // earth.rs:42-60
    Scene scene = earth_scene();
    CameraSettings camera = CameraSettings::default_().cam_pos(Vec3(5.f, 5.f, 5.f)).look_at(Vec3::zero()).field_of_view(30.f);
    Renderer renderer = Renderer::default_().width(800).height(800).samples(128).camera(camera);
    return run_example(argc, argv, "Earth", scene, renderer);
}
