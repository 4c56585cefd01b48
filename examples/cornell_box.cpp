// examples/cornell_box.rs restated against include/firework.hpp: the same calls in the same order.
#include "common.hpp"

using namespace firework;

Scene cornell_box() {   // cornell_box.rs:10-48
    Scene world = Scene::new_();

    auto red = world.add_material(LambertianMat::with_color(Vec3(0.65f, 0.05f, 0.05f)));
    auto white = world.add_material(LambertianMat::with_color(Vec3(0.73f, 0.73f, 0.73f)));
    auto green = world.add_material(LambertianMat::with_color(Vec3(0.12f, 0.45f, 0.15f)));

    auto light = world.add_material(EmissiveMat::with_color(Vec3(15.f, 15.f, 15.f)));

    world.add_object(RenderObject::new_(XZRect::new_(213.f, 343.f, 227.f, 332.f, 554.f, light)));
    world.add_object(RenderObject::new_(YZRect::new_(0.f, 555.f, 0.f, 555.f, 555.f, green)).flip_normals());
    world.add_object(RenderObject::new_(YZRect::new_(0.f, 555.f, 0.f, 555.f, 0.f, red)));
    world.add_object(RenderObject::new_(XZRect::new_(0.f, 555.f, 0.f, 555.f, 0.f, white)));
    world.add_object(RenderObject::new_(XZRect::new_(0.f, 555.f, 0.f, 555.f, 555.f, white)).flip_normals());
    world.add_object(RenderObject::new_(XYRect::new_(0.f, 555.f, 0.f, 555.f, 555.f, white)).flip_normals());
    world.add_object(RenderObject::new_(Rect3d::with_size(Vec3(165.f, 165.f, 165.f), white))
                         .rotate(Rotor3::from_rotation_xz(to_radians(18.f)))
                         .position(130.f, 0.f, 65.f));
    world.add_object(RenderObject::new_(Rect3d::with_size(Vec3(165.f, 330.f, 165.f), white))
                         .rotate(Rotor3::from_rotation_xz(to_radians(-15.f)))
                         .position(265.f, 0.f, 295.f));
    return world;
}

int main(int argc, char** argv) {   // cornell_box.rs:50-76
    Scene scene = cornell_box();
    CameraSettings camera = CameraSettings::default_().cam_pos(Vec3(278.f, 278.f, -800.f)).look_at(Vec3(278.f, 278.f, 0.f)).field_of_view(40.f);
    Renderer renderer = Renderer::default_().width(300).height(300).samples(1000).camera(camera);
    return run_example(argc, argv, "Cornell Box", scene, renderer);
}
