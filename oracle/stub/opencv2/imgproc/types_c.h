// TEST INFRASTRUCTURE ONLY: empty stand-in, see imgproc.hpp
#pragma once
