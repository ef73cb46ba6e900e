// Headless frame loop: the call pattern of the reference's Source/Main.cpp:85-206 with GLFW replaced by a
// scripted camera path (the reference moves the camera from WASD/mouse callbacks, Main.cpp:23-83).
//
//   vkrt_headless [--frames N] [--res R] [--spp S] [--depth D] [--wavefront] [--seed K] [--sleep] [--out img.ppm]
//                 [--print-camera] [--devices 0,1,...]
//
// --seed K calls srand(K) after Construct so that the seeds Draw draws from rand() are reproducible
// (the reference seeds with time(0)); --sleep keeps the 12 ms sleep of Main.cpp:195.
#include "GraphicsDevice.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

namespace
{
	Camera camera;
	FrameData frame_data;

	double now()
	{
		using clock = std::chrono::steady_clock;
		static const clock::time_point t0 = clock::now();
		return std::chrono::duration<double>(clock::now() - t0).count();
	}
}

int main(int argc, char ** argv)
{
	unsigned frames = 100, res = 1024;
	long seed = -1;
	bool do_sleep = false, print_camera = false;
	std::string out;

	for (int i = 1; i < argc; ++i)
	{
		const std::string a = argv[i];
		auto next = [&]() { return (i + 1 < argc) ? argv[++i] : "0"; };
		if (a == "--frames") frames = std::atoi(next());
		else if (a == "--res") res = std::atoi(next());
		else if (a == "--spp") GraphicsDevice::options().spp = std::atoi(next());
		else if (a == "--depth") GraphicsDevice::options().max_depth = std::atoi(next());
		else if (a == "--wavefront") GraphicsDevice::options().wavefront = true;
		else if (a == "--seed") seed = std::atol(next());
		else if (a == "--sleep") do_sleep = true;
		else if (a == "--out") out = next();
		else if (a == "--print-camera") print_camera = true;
		else if (a == "--devices")
		{
			// one Draw call per frame on several GPUs (vkrt_create_info.device_ids)
			GraphicsDevice::Options & o = GraphicsDevice::options();
			o.n_devices = 0;
			for (const char * p = next(); *p && o.n_devices < 8; )
			{
				o.devices[o.n_devices++] = std::atoi(p);
				while (*p && *p != ',') ++p;
				if (*p == ',') ++p;
			}
		}
	}

	// Main.cpp:134-139
	camera.data.pos = { 32.8509f, 30.6991f, -106.389f };
	camera.aux.pitch = 4.44998f;
	camera.aux.yaw = -602.79f;
	camera.update();

	frame_data = FrameData{};
	frame_data.light_pos = { 0.0f, 64.0f, 0.0f };          // Main.cpp:141
	frame_data.camera = camera.data;

	if (print_camera)
	{
		const float * f = &camera.data.pos.x;
		for (int k = 0; k < 16; ++k) if (k % 4 != 3) std::printf("%a ", f[k]);
		std::printf("\n");
		return 0;
	}

	GraphicsDevice device;
	{
		const GraphicsDevice::CreateInfo info{ nullptr, 3, 2, static_cast<unsigned short>(res), false };   // Main.cpp:105-115
		if (const auto r = device.Construct(info); r != GraphicsDevice::Error::SUCCESS)
		{
			std::cout << "[app] - err :: Graphics device creation failed :: " << static_cast<unsigned int>(r) << std::endl;
			return 1;
		}
	}
	if (seed >= 0) std::srand(static_cast<unsigned>(seed));

	double previous_time = now();
	unsigned frame_count = 0;

	device.Draw(frame_data);                                // Main.cpp:144

	for (unsigned f = 1; f < frames; ++f)                   // Main.cpp:148-196
	{
		const double current_time = now();
		++frame_count;
		if (current_time - previous_time >= 1.0)
		{
			std::cout << frame_count << " FPS" << std::endl;   // Main.cpp:157-163
			frame_count = 0;
			previous_time = current_time;
		}

		// scripted input in place of poll_keyboard (Main.cpp:23-60): dolly in, strafe, rise
		if (f % 3 == 0) camera.move_forward(0.5f);
		else if (f % 3 == 1) camera.move_right(0.25f);
		else camera.move_up(0.125f);

		frame_data.camera = camera.data;
		device.Draw(frame_data);

		if (do_sleep) std::this_thread::sleep_for(std::chrono::milliseconds(12));   // Main.cpp:195
	}

	device.WaitIdle();
	const double elapsed = now();
	std::cout << frames << " frames in " << elapsed << " s (" << frames / elapsed << " FPS)" << std::endl;

	if (!out.empty())
	{
		std::vector<unsigned char> img(static_cast<size_t>(res) * res * 4);
		if (!device.ReadImage(img.data(), img.size())) { std::cout << "[app] - err :: image read-back failed" << std::endl; return 1; }
		FILE * fp = std::fopen(out.c_str(), "wb");
		if (!fp) return 1;
		std::fprintf(fp, "P6\n%u %u\n255\n", res, res);
		for (int y = static_cast<int>(res) - 1; y >= 0; --y)          // row 0 is the bottom (Fullscreen.frag:16 flips it)
			for (unsigned x = 0; x < res; ++x) std::fwrite(&img[(static_cast<size_t>(y) * res + x) * 4], 1, 3, fp);
		std::fclose(fp);
	}

	device.Destruct();
	return 0;
}
