"""The drop-in boundary seen from C: include/roboy_b200.h is a plain C header (no C++, no torch types) and a C program that
only knows the header can link the library and call it.  Without a GPU the compute entry points must fail with
ROBOY_E_CUDA -- there is no CPU fallback to fall into."""
import os
import subprocess
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "roboy_b200.h")


def test_header_is_plain_c():
    for std in ("-std=c99", "-std=c11"):
        subprocess.check_call(["gcc", std, "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER])


def test_c_program_links_the_library_and_gets_an_error_code_without_a_gpu(tmp_path):
    from gym_roboy_b200 import build
    lib = build.build()
    src = tmp_path / "caller.c"
    src.write_text(textwrap.dedent("""
        #include <stdio.h>
        #include <string.h>
        #include "roboy_b200.h"
        int main(void) {
            roboy_cfg cfg;
            roboy_env *env = NULL;
            int hold_ok;
            float lo[ROBOY_DIM_ACTION], hi[ROBOY_DIM_ACTION];
            if (roboy_abi_version() != ROBOY_B200_ABI_VERSION) return 10;
            if (roboy_cfg_msj(&cfg) != ROBOY_OK) return 11;
            if (cfg.max_episode_len != 400 || cfg.angle_high <= 3.14f || cfg.act_high != 0.3f) return 12;  /* roboy_env.py:28, msj_robot.py:9,11 */
            hold_ok = roboy_hold_intervals(&cfg, lo, hi);
            if (hold_ok != ROBOY_OK || !(lo[0] < 0.0f && hi[0] > 0.0f && hi[0] < 1e-6f)) return 13;          /* simulation_client.py:38 */
            cfg.n_envs = 64;
            {
                int rc = roboy_create(&cfg, 0, &env);
                printf("roboy_create -> %d (%s)\\n", rc, roboy_last_error());
                if (rc == ROBOY_OK) { roboy_destroy(env); return 0; }   /* a GPU box */
                return rc == ROBOY_E_CUDA && env == NULL && strlen(roboy_last_error()) > 0 ? 0 : 14;
            }
        }
    """))
    exe = tmp_path / "caller"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(HEADER), "-o", str(exe), str(src),
                           lib, "-Wl,-rpath," + os.path.dirname(lib)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "roboy_create ->" in out.stdout
