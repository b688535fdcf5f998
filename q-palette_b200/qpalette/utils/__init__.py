from .matmul_had import (get_hadK, had28, matmul_hadU_cuda, matmul_hadU_head_cuda, matmul_hadUt_cuda,
                         matmul_hadUt_head_cuda)
from .mem_op import LAYER_INFO, get_dummy_quant_results, get_layer_info, get_quant_info

__all__ = ["get_hadK", "had28", "matmul_hadU_cuda", "matmul_hadU_head_cuda", "matmul_hadUt_cuda",
           "matmul_hadUt_head_cuda", "LAYER_INFO", "get_dummy_quant_results", "get_layer_info", "get_quant_info"]
