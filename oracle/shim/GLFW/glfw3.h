#pragma once
struct GLFWwindow;
