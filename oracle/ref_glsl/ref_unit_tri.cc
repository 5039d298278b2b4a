// ref_unit_tri.cc -- TEST INFRASTRUCTURE ONLY. shader/extract_tri_lights.comp (read from /root/reference, rewritten
// lexically by build_ref.py) driven like a compute dispatch: one invocation per triangle of an emissive instance.
#include "glsl_shim.hpp"
#include "ref_api.h"
#include <vector>

namespace glsl { namespace tri {
#include "ref_prelude.hpp"
void traceRayEXT(const accelerationStructureEXT&, uint, uint, uint, uint, uint, vec3, float, vec3, float, int) {}
#include "extract_tri_lights.comp.h"
}}

using namespace glsl;
using namespace glsl::tri;

extern "C" int ref_extract_tri_lights(const void* instance_table, uint32_t instance_id, const void* world_vertices, const void* index_data,
                                      uint32_t triangle_count, void* tri_lights_out)
{
    instances.o = (tri::instance*)instance_table;
    std::vector<vertex_buffer> vb(instance_id + 1);
    std::vector<index_buffer> ib(instance_id + 1);
    vb[instance_id].v = (tri::vertex*)world_vertices;
    ib[instance_id].i = (uint*)index_data;
    vertices = vb.data();
    indices = ib.data();
    tri_lights.lights = (tri::tri_light*)tri_lights_out;
    control.triangle_count = triangle_count;
    control.instance_id = instance_id;
    for(uint32_t i = 0; i < triangle_count; ++i)
    {
        gl_GlobalInvocationID = uvec3(i, 0, 0);
        extract_tri_lights_comp_main();
    }
    vertices = nullptr; indices = nullptr;
    return 0;
}
