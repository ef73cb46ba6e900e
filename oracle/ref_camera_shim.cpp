// TEST INFRASTRUCTURE.  Thin C wrapper around the REFERENCE'S OWN Source/Camera.cpp, which the
// Makefile compiles from where it lies under /root/reference (never copied into this repo).
// Used to pin the oracle's / the host mirror's camera basis (Source/Camera.cpp:48-64) and the
// default view of Source/Main.cpp:134-139.  Output goes to oracle/_ref/ only.
#include <Camera.h>
#include <GraphicsDevice.h>
#include <cstring>

static_assert(sizeof(CameraData) == 64, "CameraData");
static_assert(sizeof(FrameData) == 96, "FrameData");
static_assert(sizeof(Triangle) == 48, "Triangle");

extern "C" {
// runs Camera::update() for (pos, pitch, yaw) and writes the 64-byte CameraData
void ref_camera_update(const float pos[3], float pitch, float yaw, void *camera_data_64)
{
    Camera cam;
    cam.data.pos = glm::vec3(pos[0], pos[1], pos[2]);
    cam.aux.pitch = pitch;
    cam.aux.yaw = yaw;
    cam.update();
    std::memcpy(camera_data_64, &cam.data, 64);
}
// the FrameData Main.cpp:134-144 hands to the first Draw (aspect/seed are overwritten inside Draw)
void ref_default_frame_data(void *frame_data_96)
{
    Camera camera;
    camera.data.pos = {32.8509, 30.6991, -106.389};
    camera.aux.pitch = 4.44998;
    camera.aux.yaw = -602.79;
    camera.update();
    FrameData fd{};
    fd.light_pos = glm::vec3(0.0f, 64.0f, 0.0f);
    fd.camera = camera.data;
    std::memcpy(frame_data_96, &fd, 96);
}
unsigned ref_sizeof_frame_data() { return sizeof(FrameData); }
unsigned ref_offsetof_camera() { return offsetof(FrameData, camera); }
unsigned ref_offsetof_light_pos() { return offsetof(FrameData, light_pos); }
unsigned ref_offsetof_seed() { return offsetof(FrameData, seed); }
unsigned ref_sizeof_triangle() { return sizeof(Triangle); }
// move_* helpers (Source/Camera.cpp:12-46) for the scripted camera path of the headless loop
void ref_camera_move(void *camera_data_64, float pitch, float yaw, int op, float speed)
{
    Camera cam;
    std::memcpy(&cam.data, camera_data_64, 64);
    cam.aux.pitch = pitch; cam.aux.yaw = yaw;
    cam.update();
    switch (op) {
        case 0: cam.move_forward(speed); break;
        case 1: cam.move_backward(speed); break;
        case 2: cam.move_left(speed); break;
        case 3: cam.move_right(speed); break;
        case 4: cam.move_up(speed); break;
        case 5: cam.move_down(speed); break;
    }
    std::memcpy(camera_data_64, &cam.data, 64);
}
}
