from .abstract_matrix import AbstractDesignMatrix
from .gpu_sparse_matrix import GpuSparseDesignMatrix
from .gpu_dense_matrix import GpuDenseDesignMatrix

# The reference's class names resolve to the device-resident implementations.
SparseDesignMatrix = GpuSparseDesignMatrix
DenseDesignMatrix = GpuDenseDesignMatrix
