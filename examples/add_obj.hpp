// The `add_obj` helper of examples/suzanne.rs:15-51 (positions only) and examples/teapot.rs:17-64 (with normals, every model
// rotated): one TriangleMesh render object per model of the OBJ file.  tobj::load_obj is the library's fw_obj_load.
#pragma once
#include <string>
#include <vector>

#include "../include/firework.hpp"

inline void add_obj(firework::Scene& scene, const std::string& file_name, firework::MaterialIdx material, bool with_normals,
                    const firework::Rotor3* rotate) {
    using namespace firework;
    fw_obj* obj = nullptr;
    if (fw_obj_load(file_name.c_str(), &obj) != FW_OK) throw Error(std::string("add_obj: ") + fw_last_error());   // `assert!(obj.is_ok())`
    struct Guard { fw_obj* o; ~Guard() { fw_obj_destroy(o); } } guard{obj};
    for (int m = 0; m < fw_obj_num_models(obj); ++m) {
        uint32_t sizes[4];   // floats: positions, normals, texcoords; indices
        detail::check(fw_obj_model_sizes(obj, m, sizes), "add_obj");
        std::vector<float> pos(sizes[0]), nrm(sizes[1]), tex(sizes[2]);
        std::vector<uint32_t> idx(sizes[3]);
        detail::check(fw_obj_model_copy(obj, m, pos.data(), nrm.data(), tex.data(), idx.data()), "add_obj");
        std::vector<Vec3> verts, normals;
        for (size_t i = 0; i + 2 < pos.size(); i += 3) verts.push_back(Vec3(pos[i], pos[i + 1], pos[i + 2]));
        if (with_normals)
            for (size_t i = 0; i + 2 < nrm.size(); i += 3) normals.push_back(Vec3(nrm[i], nrm[i + 1], nrm[i + 2]));
        RenderObject ro = RenderObject::new_(TriangleMesh::new_(std::move(verts), std::move(idx), std::move(normals), {}, material));
        scene.add_object(rotate ? std::move(ro).rotate(*rotate) : std::move(ro));
    }
}
