"""Import shim: the package directory is ``genome-assembly_b200/`` (not a valid Python identifier),
so this module turns itself into that package: ``import genome_assembly_b200`` then behaves as if the
directory were importable, including ``genome_assembly_b200.synth`` etc."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "genome-assembly_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
_init = _os.path.join(__path__[0], "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
