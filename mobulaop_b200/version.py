__version__ = '0.1.0'
# the reference release whose ROIAlign interface this package mirrors
REFERENCE_VERSION = '2.51'
