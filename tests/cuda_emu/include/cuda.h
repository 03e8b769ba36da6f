// tests/cuda_emu: empty stand-in for the CUDA driver header (TEST INFRASTRUCTURE ONLY)
#pragma once
