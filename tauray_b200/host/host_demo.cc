// host_demo.cc -- drives the C++ host mirror (tauray_b200.hh) the way Tauray's replay_viewer drives
// rt_renderer (src/tauray.cc:1065-1117): build renderer, set scene, per frame reset_accumulation + render,
// read the image back. Used by tests/test_gpu_host_mirror.py: reads a scene dump, writes the raw rgba32f image.
//
//   host_demo <scene.bin> <out.raw> <width> <height> <spp> <depth> <devices>
//
// scene.bin layout (little endian): u32 instance_count, then per instance {tr_instance, u32 vertex_count,
// u32 index_count, u32 opaque, vertices, indices}; u32 directional_light_count + entries; u32 point_light_count
// + entries; u32 view_count + tr_camera_pair entries. (No textures.)
#include "tauray_b200.hh"
#include <cstdio>
#include <cstdlib>
#include <fstream>

using namespace tr_b200;

template<typename T> static void rd(std::ifstream& f, T* dst, size_t n) { f.read(reinterpret_cast<char*>(dst), sizeof(T) * n); }

int main(int argc, char** argv)
{
    if(argc < 8) { std::fprintf(stderr, "usage: host_demo scene.bin out.raw w h spp depth devices\n"); return 2; }
    try
    {
        std::ifstream f(argv[1], std::ios::binary);
        if(!f) throw std::runtime_error("cannot open scene dump");
        uint32_t n = 0;
        rd(f, &n, 1);
        std::vector<tr_instance> instances(n);
        std::vector<tr_mesh_ref> meshes(n);
        std::vector<std::vector<tr_vertex>> verts(n);
        std::vector<std::vector<uint32_t>> idx(n);
        for(uint32_t i = 0; i < n; ++i)
        {
            uint32_t vc, ic, opaque;
            rd(f, &instances[i], 1); rd(f, &vc, 1); rd(f, &ic, 1); rd(f, &opaque, 1);
            verts[i].resize(vc); idx[i].resize(ic);
            rd(f, verts[i].data(), vc); rd(f, idx[i].data(), ic);
            meshes[i] = tr_mesh_ref{verts[i].data(), idx[i].data(), vc, ic, opaque, 0};
        }
        uint32_t ndl = 0, npl = 0, nv = 0;
        rd(f, &ndl, 1); std::vector<tr_directional_light> dl(ndl); rd(f, dl.data(), ndl);
        rd(f, &npl, 1); std::vector<tr_point_light> pl(npl); rd(f, pl.data(), npl);
        rd(f, &nv, 1); std::vector<tr_camera_pair> cams(nv); rd(f, cams.data(), nv);

        tr_pt_scene_desc desc;
        std::memset(&desc, 0, sizeof(desc));
        desc.instance_count = n; desc.instances = instances.data(); desc.meshes = meshes.data();
        desc.directional_light_count = ndl; desc.directional_lights = dl.data();
        desc.point_light_count = npl; desc.point_lights = pl.data();
        desc.environment_proj = -1;

        rt_renderer::options opt;
        // what create_renderer copies from the CLI options (src/tauray.cc:392-403,472-488)
        opt.max_ray_depth = std::atoi(argv[6]);
        opt.min_ray_dist = 1e-4f;
        opt.samples_per_pixel = std::atoi(argv[5]);
        opt.samples_per_pass = 1;
        opt.local_sampler = path_tracer_stage::sampler_type::SOBOL_OWEN;
        opt.film = film_filter::POINT;
        opt.film_radius = 0.5f;
        opt.tri_light_mode = tri_light_sampling_mode::SOLID_ANGLE;
        const int ndev = std::atoi(argv[7]);
        std::vector<int> devices;
        for(int i = 0; i < ndev; ++i) devices.push_back(i);   // with one physical GPU use `devices = {0, 0}` (== --fake-devices)
        if(const char* fake = std::getenv("TR_FAKE_DEVICES")) { devices.assign((size_t)std::atoi(fake), 0); }

        uvec2 size{(unsigned)std::atoi(argv[3]), (unsigned)std::atoi(argv[4])};
        const std::string mode = argc > 8 ? argv[8] : "";
        if(mode == "direct")
        {
            // --renderer=direct: direct_stage (src/direct_stage.hh) on one device, 2 NEE samples per pass
            direct_stage::options dopt;
            dopt.max_ray_depth = opt.max_ray_depth; dopt.min_ray_dist = opt.min_ray_dist;
            dopt.samples_per_pixel = opt.samples_per_pixel; dopt.samples_per_pass = 2;
            dopt.local_sampler = opt.local_sampler;
            dopt.film = film_filter::BOX; dopt.film_radius = 0.5f;
            dopt.tri_light_mode = tri_light_sampling_mode::SOLID_ANGLE;
            dopt.distribution = get_device_distribution_params(size, DISTRIBUTION_SHUFFLED_STRIPS, 0.0, 1.0, 0, 1, true);
            direct_stage stage(devices[0], desc, cams.data(), nv, dopt);
            stage.run();
            stage.sync();
            std::vector<float> img = stage.read_color();
            std::ofstream o(argv[2], std::ios::binary);
            o.write(reinterpret_cast<const char*>(img.data()), img.size() * sizeof(float));
            std::printf("host_demo: direct, %zu floats\n", img.size());
            return 0;
        }
        const bool rebalance = mode == "rebalance";
        if(mode == "scanline") opt.strategy = DISTRIBUTION_SCANLINE;   // --distribution-strategy=scanline across the devices
        opt.accumulate = rebalance;
        rt_renderer rr(devices, size, opt);
        rr.set_scene(desc, cams.data(), nv);
        rr.reset_accumulation();
        rr.render();
        if(rebalance)
        {
            // a workload change under accumulation (src/rt_renderer.cc:136-183): 2 frames, new strips, 1 frame
            rr.render();
            std::vector<double> w(rr.device_count(), rr.device_count() > 1 ? 0.7 / (double)(rr.device_count() - 1) : 1.0);
            if(rr.device_count() > 1) w[0] = 0.3;
            rr.set_device_workloads(w);
            rr.render();
        }
        if(mode == "animate" || mode == "refit")
        {
            if(mode == "refit") rr.set_update_mode(TR_UPDATE_REFIT);
            // scene_stage::update of an animated scene: instance 6 slides by (0.2, 0, 0.1); the previous transform
            // feeds the screen-motion layer. Pure translation: model_normal is unchanged.
            if(n < 7) throw std::runtime_error("animate: the scene needs at least 7 instances");
            std::vector<tr_instance> moved = instances;
            std::memcpy(moved[6].model_prev, moved[6].model, sizeof(moved[6].model));
            moved[6].model[12] += 0.2f; moved[6].model[14] += 0.1f;
            rr.update_instances(moved.data(), n);
            rr.reset_accumulation();
            rr.render();
        }
        std::vector<float> img = rr.read_image();
        std::ofstream o(argv[2], std::ios::binary);
        o.write(reinterpret_cast<const char*>(img.data()), img.size() * sizeof(float));
        load_balancer lb(rr.device_count());
        lb.update(rr.device_times_ms());
        std::printf("host_demo: %zu devices, %zu floats, device0 %.3f ms, workloads[0] %.4f\n",
                    rr.device_count(), img.size(), rr.device_times_ms()[0], lb.workloads[0]);
    }
    catch(const std::runtime_error& e)
    {
        std::fprintf(stderr, "host_demo: %s\n", e.what());   // src/main.cc:36-42: runtime_error -> exit 1
        return 1;
    }
    return 0;
}
