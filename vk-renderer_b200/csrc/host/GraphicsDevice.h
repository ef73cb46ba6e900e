// Host-side drop-in for the reference's engine interface on the ray-tracing path, in C++.
//
// Same type names, member names, argument meaning and error values as the reference's
// Include/GraphicsDevice.h:13-97 and Include/Camera.h:5-43, so that Source/Main.cpp's call pattern
// (:105-119 Construct, :134-144 first Draw, :148-196 loop, :200-202 WaitIdle/Destruct) compiles against this
// header unchanged -- but implemented on libvkrt_cuda (include/vkrt.h) instead of Vulkan.  Written without glm:
// the PODs are plain floats with the reference's 16-byte alignment, byte-identical to the shader blocks.
// There is no window or swapchain here: CreateInfo::window is ignored and the "swapchain
// extent" that Draw turns into aspect_ratio (Source/GraphicsDevice.cpp:1260) is 1024x768 unless set.
#pragma once
#include <cstdint>

struct GLFWwindow;

struct alignas(16) Vec3 { float x, y, z; };        // a glm::vec3 with alignas(16): 12 B of data in a 16 B slot

struct CameraData      // ref: Include/Camera.h:5-12
{
	Vec3 pos;
	Vec3 dir;
	Vec3 right;
	Vec3 up;
};

struct CameraDataAux   // ref: Include/Camera.h:14-20
{
	float front[3];
	float pitch;
	float yaw;
};

struct Camera          // ref: Include/Camera.h:22-43, Source/Camera.cpp
{
	CameraData    data;
	CameraDataAux aux;

	Camera();

	void move_forward(float speed);
	void move_backward(float speed);
	void move_left(float speed);
	void move_right(float speed);
	void move_up(float speed);
	void move_down(float speed);

	void update();
};

struct Triangle        // ref: Include/GraphicsDevice.h:13-18
{
	Vec3 v0;
	Vec3 v1;
	Vec3 v2;
};

struct FrameData       // ref: Include/GraphicsDevice.h:20-29
{
	alignas(4) float aspect_ratio;
	alignas(4) float seed;
	alignas(16) Vec3 light_pos;
	alignas(16) CameraData camera;
};

static_assert(sizeof(CameraData) == 64 && sizeof(FrameData) == 96 && sizeof(Triangle) == 48, "layouts of the reference");

struct GraphicsDevice final
{
	enum class Error : signed char      // ref: Include/GraphicsDevice.h:46-52
	{
		SUCCESS,
		NO_SUITABLE_GPU,
		NO_SUITABLE_SURFACE,
		UNKNOWN
	};

	struct CreateInfo final             // ref: Include/GraphicsDevice.h:57-72
	{
		GLFWwindow * window;
		unsigned char swapchainSize;
		unsigned char framesInFlight;
		unsigned short raytrace_resolution;
		bool debug;
	};

	Error Construct(const CreateInfo & info);
	Error Destruct();
	void Draw(const FrameData & frame_data);
	void WaitIdle();

	// --- headless extras (not in the reference) -------------------------------------------------------
	// the shader's compile-time constants / the swapchain extent, to be set before Construct
	// n_devices > 1: the ONE Draw call renders on devices[0 .. n_devices) (vkrt_create_info.device_ids, tile shards
	// exchanged over peer memory into devices[0]) -- the engine code does not change
	struct Options { unsigned spp = 4, max_depth = 4, extent_w = 1024, extent_h = 768; int device = 0; bool wavefront = false;
	                 unsigned n_devices = 0; int devices[8] = {0, 1, 2, 3, 4, 5, 6, 7}; };
	static Options & options();
	// copies the most recent traced image (rgba8, row 0 = bottom like the shader's imageStore) to the host
	bool ReadImage(unsigned char * rgba8, unsigned long long bytes);
	unsigned long long frames_drawn() const;
};
