// GraphicsDevice on libvkrt_cuda: what Source/GraphicsDevice.cpp does for the compute path, without Vulkan.
//   Construct  -> vkrt_create + the shader-constant scene of Tracer.comp (the shader the host loads, :1091)
//   Draw       -> copy FrameData, overwrite aspect_ratio and seed like :1258-1262, vkrt_draw (asynchronous, two
//                 frames in flight like FRAMES_IN_FLIGHT)
//   WaitIdle   -> vkrt_wait_idle;  Destruct -> vkrt_destroy
// Like the reference the state is a file-local singleton (Source/GraphicsDevice.cpp:40-43) and failures print
// "[app] - err :: ..." to std::cout.
#include "GraphicsDevice.h"

#include <vkrt.h>

#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>

static_assert(sizeof(FrameData) == sizeof(vkrt_frame_data), "FrameData is the 96-byte push-constant block");

namespace
{
	struct State
	{
		vkrt_ctx * ctx = nullptr;
		unsigned extent_w = 1024, extent_h = 768;
		unsigned long long frames = 0;
		unsigned resolution = 0;
	} state;
}

GraphicsDevice::Options & GraphicsDevice::options()
{
	static Options o;
	return o;
}

GraphicsDevice::Error GraphicsDevice::Construct(const CreateInfo & info)
{
	const Options & o = options();

	vkrt_create_info ci;
	std::memset(&ci, 0, sizeof ci);
	ci.struct_size      = sizeof ci;
	ci.width            = info.raytrace_resolution;          // square target, Source/GraphicsDevice.cpp:669-699
	ci.height           = info.raytrace_resolution;
	ci.spp              = o.spp;                             // SAMPLES, Tracer.comp:180
	ci.max_depth        = o.max_depth;                       // DEPTH,   Tracer.comp:179
	ci.integrator       = VKRT_INTEGRATOR_PATH;
	ci.variant          = o.wavefront ? VKRT_VARIANT_WAVEFRONT : VKRT_VARIANT_MEGAKERNEL;
	ci.frames_in_flight = info.framesInFlight;
	ci.device_id        = o.device;
	ci.n_devices        = o.n_devices;                       // > 1: one Draw, several GPUs (the reference is single-GPU)
	for (unsigned i = 0; i < 8; ++i) ci.device_ids[i] = o.devices[i];

	if (const vkrt_error e = vkrt_create(&ci, &state.ctx); e != VKRT_SUCCESS)
	{
		std::cout << vkrt_last_error_string(nullptr) << std::endl;
		return static_cast<Error>(e <= 3 ? e : 3);           // the C ABI's extra codes fold into UNKNOWN
	}

	if (const vkrt_error e = vkrt_use_default_scene(state.ctx, VKRT_SCENE_TRACER); e != VKRT_SUCCESS)
	{
		std::cout << vkrt_last_error_string(state.ctx) << std::endl;
		return Error::UNKNOWN;
	}

	state.extent_w = o.extent_w;
	state.extent_h = o.extent_h;
	state.resolution = info.raytrace_resolution;
	state.frames = 0;

	std::srand(static_cast<unsigned>(std::time(nullptr)));   // Source/GraphicsDevice.cpp:1133

	return Error::SUCCESS;
}

GraphicsDevice::Error GraphicsDevice::Destruct()
{
	if (state.ctx == nullptr) return Error::UNKNOWN;
	vkrt_destroy(state.ctx);
	state.ctx = nullptr;
	return Error::SUCCESS;
}

void GraphicsDevice::Draw(const FrameData & frame_data)
{
	FrameData frame_data_real = frame_data;                                                              // :1258
	frame_data_real.aspect_ratio = static_cast<float>(state.extent_w) / static_cast<float>(state.extent_h);   // :1260
	frame_data_real.seed = static_cast<float>(std::rand()) / static_cast<float>(RAND_MAX);              // :1262

	// void like the reference, which ignores every VkResult in Draw; an error is at least reported
	if (vkrt_draw(state.ctx, reinterpret_cast<const vkrt_frame_data *>(&frame_data_real)) != VKRT_SUCCESS)
		std::cout << vkrt_last_error_string(state.ctx) << std::endl;
	++state.frames;
}

void GraphicsDevice::WaitIdle()
{
	vkrt_wait_idle(state.ctx);
}

bool GraphicsDevice::ReadImage(unsigned char * rgba8, unsigned long long bytes)
{
	return vkrt_read_rgba8(state.ctx, rgba8, bytes) == VKRT_SUCCESS;
}

unsigned long long GraphicsDevice::frames_drawn() const
{
	return state.frames;
}
