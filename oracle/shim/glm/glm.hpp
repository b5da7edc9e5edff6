// Minimal stand-in for the subset of GLM the reference hot path uses (glm is not installed here).
// Arithmetic order mirrors scalar GLM: component-wise ops, dot = x*x + y*y + z*z left to right,
// normalize = v * (1/sqrt(dot)). Test infrastructure only (oracle/_ref build).
#pragma once
#include <cmath>
namespace glm {
template<typename T> struct tvec3 {
	T x, y, z;
	tvec3() {}
	tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
	explicit tvec3(T a) : x(a), y(a), z(a) {}
	template<typename U> explicit tvec3(const tvec3<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
	T& operator[](int i) { return (&x)[i]; }
	const T& operator[](int i) const { return (&x)[i]; }
	tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
	tvec3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
	tvec3& operator/=(T s) { x /= s; y /= s; z /= s; return *this; }
};
template<typename T> tvec3<T> operator+(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template<typename T> tvec3<T> operator-(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template<typename T> tvec3<T> operator*(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template<typename T> tvec3<T> operator/(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x / b.x, a.y / b.y, a.z / b.z); }
template<typename T> tvec3<T> operator-(const tvec3<T>& a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template<typename T> tvec3<T> operator+(const tvec3<T>& a, T s) { return tvec3<T>(a.x + s, a.y + s, a.z + s); }
template<typename T> tvec3<T> operator-(const tvec3<T>& a, T s) { return tvec3<T>(a.x - s, a.y - s, a.z - s); }
template<typename T> tvec3<T> operator*(const tvec3<T>& a, T s) { return tvec3<T>(a.x * s, a.y * s, a.z * s); }
template<typename T> tvec3<T> operator*(T s, const tvec3<T>& a) { return tvec3<T>(a.x * s, a.y * s, a.z * s); }
template<typename T> tvec3<T> operator/(const tvec3<T>& a, T s) { return tvec3<T>(a.x / s, a.y / s, a.z / s); }
typedef tvec3<float> vec3;
typedef tvec3<int> ivec3;
inline float dot(const vec3& a, const vec3& b) { vec3 t = a * b; return t.x + t.y + t.z; }
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline float distance(const vec3& a, const vec3& b) { return length(b - a); }
inline vec3 normalize(const vec3& a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline vec3 cross(const vec3& x, const vec3& y) { return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
using std::isnan; // same entity as ::isnan from <math.h>, so `using namespace glm` stays unambiguous
template<typename T> T min(T a, T b) { return b < a ? b : a; }
template<typename T> T max(T a, T b) { return a < b ? b : a; }
}
