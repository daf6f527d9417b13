"""The Lennard-Jones kernel derives five divisions per pair from ONE correctly rounded reciprocal
(csrc/objectives.cu: div_by).  tests/c/div_check.c replays the same three-instruction sequence on the CPU (gcc,
hardware FMA) against the IEEE quotient on 4e7 random and adversarial operand pairs: they must never differ."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shared_reciprocal_division_is_correctly_rounded():
    exe = os.path.join(ROOT, "build", "div_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    try:
        hw_fma = " fma " in open("/proc/cpuinfo").read()
    except OSError:
        hw_fma = False
    flags = ["-mfma"] if hw_fma else []       # without it fma() is libm's exact software FMA (slower, same result)
    r = subprocess.run(["gcc", "-O2", "-ffp-contract=off", *flags, "-o", exe, os.path.join(ROOT, "tests", "c", "div_check.c"), "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "mismatches=0 " in r.stdout, r.stdout
