// dist.cu -- placeholder, replaced below
