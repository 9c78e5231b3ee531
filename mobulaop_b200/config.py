"""Global settings object, same contract as the reference's `mobula.config`
(/root/reference/mobula/config.py:27-66): attributes are fixed at creation, assignments
are type checked, `TempConfig(**kw)` scopes a change.

Only the flags the ROIAlign path reads do anything here; the reference's build flags are
kept as inert attributes so user code that sets them keeps running.
"""

_LIVE = dict(
    # backward algorithm: False = fastest (rounding order not reproducible),
    # True = bit-exact canonical order (== single-thread reference)
    ROIALIGN_DETERMINISTIC=False,
    # kernel family: 'auto' | 'gather' | 'plane' | 'resident' | 'sweep' | 'pull'
    ROIALIGN_ALGO='auto',
    # the reference dispatches on this (func.py:143); only 'cuda' exists here
    GPU_BACKEND='cuda',
    # reference: push kernels to MXNet's engine (func.py:203); here: launch on the
    # caller's current CUDA stream without synchronising
    USING_ASYNC_EXEC=True,
)

_INERT = dict(
    TARGET='mobula_op', BUILD_PATH='./', BUILD_IN_LOCAL_PATH=True, SHOW_BUILDING_COMMAND=False,
    MAX_BUILDING_WORKER_NUM=8, DEBUG=False, USING_OPENMP=True, USING_CBLAS=False,
    HOST_NUM_THREADS=0, USING_HIGH_LEVEL_WARNINGS=False, USING_OPTIMIZATION=True,
    CXX='g++', NVCC='nvcc', HIPCC='hipcc', CUDA_DIR='/usr/local/cuda', HIP_DIR='/opt/rocm/hip',
)


class Config(object):
    def __init__(self):
        values = dict(_INERT)
        values.update(_LIVE)
        object.__setattr__(self, '_values', values)

    def __getattr__(self, name):
        try:
            return object.__getattribute__(self, '_values')[name]
        except KeyError:
            raise AttributeError("Config has no attribute '%s'" % name)

    def __setattr__(self, name, value):
        values = object.__getattribute__(self, '_values')
        if name not in values:
            raise AttributeError("Config has no attribute '%s'" % name)
        want, got = type(values[name]), type(value)
        if want is not got:
            raise TypeError('The type of config attribute `%s` is not consistent, '
                            'target %s vs value %s.' % (name, want, got))
        if name == 'GPU_BACKEND' and value != 'cuda':
            raise ValueError("only the 'cuda' (sm_100a) backend exists in this build")
        if name == 'ROIALIGN_ALGO' and value not in ('auto', 'gather', 'plane', 'resident', 'sweep', 'pull'):
            raise ValueError('unknown ROIAlign algorithm %r' % (value,))
        values[name] = value

    def __dir__(self):
        return sorted(object.__getattribute__(self, '_values'))


config = Config()


class TempConfig(object):
    """`with mobula.config.TempConfig(ROIALIGN_DETERMINISTIC=True): ...`"""

    def __init__(self, **kwargs):
        self._new = kwargs
        self._old = {}

    def __enter__(self):
        for key in self._new:
            if key not in dir(config):
                raise AttributeError("'mobula.config' object has no attribute '%s'" % key)
        for key, value in self._new.items():
            self._old[key] = getattr(config, key)
            setattr(config, key, value)
        return config

    def __exit__(self, *exc):
        for key, value in self._old.items():
            setattr(config, key, value)
        return False


Config.TempConfig = TempConfig
