# Container image for the legacy one-rank-per-container deployment (reference Dockerfile:1-14).
# Build on a CUDA 12.9 / PyTorch >= 2.11 base; the extension is compiled in-tree for sm_100a.
FROM pytorch/pytorch:latest
WORKDIR /app
COPY . /app
RUN pip install --no-cache-dir pandas psutil numpy && \
    (python -c "import __graft_entry__ as g; g.build()" || echo "CUDA extension not built (no nvcc); torch backend only")
ENV PYTHONUNBUFFERED=1
CMD ["python", "train.py"]
