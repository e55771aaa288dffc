def simd_available() -> bool:
    """True when the host CPU has what the wide (AVX-512) forms of the host lanes need (csrc/simd_text.hpp); the scalar forms
    give the same results everywhere."""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return all(f in flags for f in ("avx512f", "avx512bw", "avx512vl", "avx512_vbmi2", "bmi2", "pclmulqdq"))
