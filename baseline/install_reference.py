"""Puts the UNMODIFIED reference (hkchengrex/Cutie, read-only at /root/reference) under baseline/_ref/ so that it
travels to the GPU box (baseline/_ref/ is git-ignored -- never part of this repo's history -- but not gpurun-ignored).

Used by: bench.py --impl reference (the reference's own implementation timed on the box's host cores) and the GPU-side
reference comparator of tests/ (tests/ref_runner.py: the reference run in eager fp32 on the same GPU).  Never imported
by the product.

Recipe: `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref <copy of the
reference>` first.  In this image that fails (build backend `hatchling` is not installed and there is no index); the
fallback installs exactly what the reference's wheel would contain -- pyproject.toml: `[tool.hatch.build.targets.wheel]
packages = ["cutie"]` -- i.e. the `cutie/` package directory, byte for byte.  The outcome is written to
baseline/_ref/INSTALL.txt."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SRC = os.environ.get('CUTIE_REFERENCE_SRC', '/root/reference')


def installed() -> bool:
    return os.path.isfile(os.path.join(DST, 'cutie', 'inference', 'inference_core.py'))


def install(force: bool = False, try_pip: bool = True) -> str:
    if installed() and not force:
        return 'present'
    if not os.path.isdir(os.path.join(SRC, 'cutie')):
        return 'no reference tree at ' + SRC
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    note = ''
    if try_pip:
        with tempfile.TemporaryDirectory() as tmp:
            cp = os.path.join(tmp, 'reference')
            shutil.copytree(SRC, cp, ignore=shutil.ignore_patterns('.git', 'docs', 'examples'))
            r = subprocess.run([sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps',
                                '--find-links', '/opt/wheelhouse', '--target', DST, cp],
                               capture_output=True, text=True)
        if r.returncode == 0 and installed():
            note = 'pip install --target baseline/_ref: ok'
        else:
            tail = (r.stderr or r.stdout).strip().splitlines()[-1:] or ['?']
            note = 'pip install failed (' + tail[0][:160] + '); '
    if not installed():
        shutil.rmtree(DST, ignore_errors=True)
        os.makedirs(DST, exist_ok=True)
        shutil.copytree(os.path.join(SRC, 'cutie'), os.path.join(DST, 'cutie'),
                        ignore=shutil.ignore_patterns('__pycache__'))
        note += 'installed the wheel content (packages = ["cutie"]) by directory copy'
    with open(os.path.join(DST, 'INSTALL.txt'), 'w') as f:
        f.write(note + '\n')
    return note


if __name__ == '__main__':
    print(install(force='--force' in sys.argv))
