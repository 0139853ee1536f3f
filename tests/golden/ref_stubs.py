"""Module stubs that let parts of the UNMODIFIED reference (/root/reference) import in the authoring container.

Used only by the golden-vector generators under tests/golden/ (the GPU box has no /root/reference).  The reference's
GPflow path needs gpflow / tensorflow / tensorflow_probability, and its orchestrator imports xarray, PyTables,
matplotlib, seaborn, cartopy, pyproj and dataclasses_json at module level; none are installed here.  The stubs
below provide exactly the attributes those imports touch (SURVEY.md section 8c "stub recipe"); nothing numerical is
stubbed -- DataLoader (scipy KDTree), PredictionLocations (numba), utils' table shapers and LocalExpertOI.run itself
execute as written.
"""
import sys
import types

REF = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs(orchestrator=False):
    tf = _mod("tensorflow")
    tfp_ = _mod("tensorflow.python")
    tfc = _mod("tensorflow.python.client")
    dl = _mod("tensorflow.python.client.device_lib", list_local_devices=lambda: [])
    tf.python, tfp_.client, tfc.device_lib = tfp_, tfc, dl
    tb = _mod("tables")
    tbe = _mod("tables.exceptions", HDF5ExtError=type("HDF5ExtError", (Exception,), {}))
    tb.exceptions = tbe

    class _DA:  # xarray.DataArray / Dataset placeholders (isinstance checks only)
        pass

    class _DS:
        pass

    xr = _mod("xarray", DataArray=_DA, Dataset=_DS)
    xc = _mod("xarray.core")
    xd = _mod("xarray.core.dataarray", DataArray=_DA, Dataset=_DS)
    xr.core, xc.dataarray = xc, xd
    _mod("pyproj", Transformer=object)
    if orchestrator:     # GPSat.local_experts / plot_utils / config_dataclasses module-level imports
        mpl = _mod("matplotlib")
        plt = _mod("matplotlib.pyplot")
        mb = _mod("matplotlib.backends")
        mbp = _mod("matplotlib.backends.backend_pdf", PdfPages=object)
        mpl.pyplot, mpl.backends, mb.backend_pdf = plt, mb, mbp
        _mod("seaborn")
        _mod("dataclasses_json", dataclass_json=lambda cls: cls, config=lambda **kw: {})
    if REF not in sys.path:
        sys.path.insert(0, REF)
