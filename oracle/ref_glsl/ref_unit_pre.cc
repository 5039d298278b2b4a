// ref_unit_pre.cc -- TEST INFRASTRUCTURE ONLY. shader/pre_transform.comp (read from /root/reference, rewritten lexically
// by build_ref.py) driven like a compute dispatch: one invocation per vertex.
#include "glsl_shim.hpp"
#include "ref_api.h"

namespace glsl { namespace pre {
#include "ref_prelude.hpp"
void traceRayEXT(const accelerationStructureEXT&, uint, uint, uint, uint, uint, vec3, float, vec3, float, int) {}
#include "pre_transform.comp.h"
}}

using namespace glsl;
using namespace glsl::pre;

extern "C" int ref_pre_transform(const void* instance_record, const void* vertices_in, void* vertices_out, uint32_t vertex_count)
{
    instances.o = (pre::instance*)instance_record;
    input_verts.v = (pre::vertex*)vertices_in;
    output_verts.v = (pre::vertex*)vertices_out;
    control.vertex_count = vertex_count;
    control.instance_id = 0;
    for(uint32_t i = 0; i < vertex_count; ++i)
    {
        gl_GlobalInvocationID = uvec3(i, 0, 0);
        pre_transform_comp_main();
    }
    return 0;
}
