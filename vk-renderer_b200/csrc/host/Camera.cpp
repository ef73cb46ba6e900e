// Camera of the reference (Source/Camera.cpp) restated without glm.  Arithmetic notes that matter for
// bit-exactness against the reference's own build (tests/golden/camera_poses.json):
//   * glm::radians(x) = x * 0.01745329251994329576923690768489f, in binary32;
//   * the reference calls unqualified cos()/sin() on floats, which resolves to the C double functions, so the
//     products front.x/.z are formed in binary64 and rounded once on assignment (Camera.cpp:54-56);
//   * glm::normalize(v) = v * (1 / sqrt(dot(v, v))) with dot summed (x*x + y*y) + z*z in binary32;
//   * glm::cross(a, b) = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y).
// Build with -ffp-contract=off so that the compiler fuses nothing.
#include "GraphicsDevice.h"

#include <algorithm>
#include <cmath>

namespace
{
	struct F3 { float x, y, z; };

	F3 normalize(F3 v)
	{
		const float d = (v.x * v.x + v.y * v.y) + v.z * v.z;
		const float inv = 1.0f / std::sqrt(d);
		return { v.x * inv, v.y * inv, v.z * inv };
	}

	F3 cross(F3 a, F3 b)
	{
		return { a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y };
	}

	void step(Camera & cam, F3 dir, float speed)
	{
		cam.data.pos.x += dir.x * speed;
		cam.data.pos.y += dir.y * speed;
		cam.data.pos.z += dir.z * speed;
		cam.update();
	}
}

Camera::Camera()
	: data{}
	, aux{ { 0.0f, 0.0f, -1.0f }, 0.0f, 90.0f }      // Camera.cpp:5-7
{
	data.pos = { 0.0f, 64.0f, -48.0f };
	update();
}

void Camera::move_forward(float speed)  { step(*this, {  aux.front[0],  aux.front[1],  aux.front[2] }, speed); }
void Camera::move_backward(float speed) { step(*this, { -aux.front[0], -aux.front[1], -aux.front[2] }, speed); }
void Camera::move_left(float speed)     { step(*this, { -data.right.x, -data.right.y, -data.right.z }, speed); }
void Camera::move_right(float speed)    { step(*this, {  data.right.x,  data.right.y,  data.right.z }, speed); }
void Camera::move_up(float speed)       { step(*this, {  data.up.x,  data.up.y,  data.up.z }, speed); }
void Camera::move_down(float speed)     { step(*this, { -data.up.x, -data.up.y, -data.up.z }, speed); }

void Camera::update()                               // Camera.cpp:48-64
{
	aux.pitch = std::clamp(aux.pitch, -89.0f, 89.0f);

	const float rad = 0.01745329251994329576923690768489f;
	const double p = static_cast<double>(aux.pitch * rad), y = static_cast<double>(aux.yaw * rad);

	const F3 front = normalize({ static_cast<float>(std::cos(p) * std::cos(y)),
	                             static_cast<float>(std::sin(p)),
	                             static_cast<float>(std::cos(p) * std::sin(y)) });
	aux.front[0] = front.x; aux.front[1] = front.y; aux.front[2] = front.z;

	const F3 pos = { data.pos.x, data.pos.y, data.pos.z };
	const F3 dir = normalize({ (pos.x + front.x) - pos.x, (pos.y + front.y) - pos.y, (pos.z + front.z) - pos.z });
	const F3 right = normalize(cross({ 0.0f, 1.0f, 0.0f }, dir));
	const F3 up = cross(dir, right);

	data.dir   = { dir.x, dir.y, dir.z };
	data.right = { right.x, right.y, right.z };
	data.up    = { up.x, up.y, up.z };
}
