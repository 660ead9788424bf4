# Jupyter inside the control container (reference: control/Docker/jupyter_notebook_config.py — port 9999, no browser)
c = get_config()  # noqa: F821
c.NotebookApp.ip = "0.0.0.0"
c.NotebookApp.port = 9999
c.NotebookApp.open_browser = False
c.NotebookApp.allow_root = True
