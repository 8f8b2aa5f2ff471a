# Build / run image for megatronapp_b200 on B200 (sm_100a).  The reference's Dockerfile builds two RDMA
# C++ extensions against rdma-core; here the single native module is compiled in-tree by nvcc.
FROM nvcr.io/nvidia/pytorch:25.04-py3
WORKDIR /workspace/megatronapp_b200
COPY . .
RUN pip install --no-cache-dir pybind11 websockets sentencepiece pyyaml
RUN python -c "import __graft_entry__ as g; g.build()"
ENV PYTHONPATH=/workspace/megatronapp_b200:/workspace/megatronapp_b200/compat
CMD ["python", "-m", "pytest", "tests", "-q", "-m", "not gpu"]
