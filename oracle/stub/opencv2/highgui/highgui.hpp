// TEST INFRASTRUCTURE ONLY: empty stand-in, see ../imgproc/imgproc.hpp
#pragma once
