"""Host logic of the driver mirrors (harness.py): ImageFolder indexing against torchvision's own class (18:35)."""
import numpy as np
import pytest
from PIL import Image


def _tree(tmp_path):
    rng = np.random.default_rng(0)
    root = tmp_path / "data"
    layout = {"00010": ["b.png", "a.ppm", "notes.txt", "sub/c.PNG"], "00002": ["z.png"], "00001": ["k.jpg", "k.json"]}
    for cls, names in layout.items():
        for nm in names:
            f = root / cls / nm
            f.parent.mkdir(parents=True, exist_ok=True)
            if f.suffix.lower() in (".png", ".ppm", ".jpg"):
                Image.fromarray(rng.integers(0, 256, (9, 7, 3), dtype=np.uint8)).save(f)
            else:
                f.write_text("x")
    (root / "stray.png").write_bytes(b"")          # files directly under the root are not samples
    return root


def test_image_folder_matches_torchvision(tmp_path):
    from torchvision import datasets
    from b200restore import harness
    root = _tree(tmp_path)
    ref = datasets.ImageFolder(str(root))
    samples, classes = harness.image_folder(root)
    assert classes == ref.classes == ["00001", "00002", "00010"]
    assert samples == ref.samples
    assert harness.IMG_EXTENSIONS == tuple(datasets.folder.IMG_EXTENSIONS)


def test_image_folder_errors_like_torchvision(tmp_path):
    from torchvision import datasets
    from b200restore import harness
    empty = tmp_path / "empty"
    empty.mkdir()
    with pytest.raises(FileNotFoundError):
        datasets.ImageFolder(str(empty))
    with pytest.raises(FileNotFoundError):
        harness.image_folder(empty)
    (empty / "00000").mkdir()
    (empty / "00000" / "readme.txt").write_text("x")
    with pytest.raises(FileNotFoundError):
        datasets.ImageFolder(str(empty))
    with pytest.raises(FileNotFoundError):
        harness.image_folder(empty)


def test_missing_directory_is_skipped_like_the_reference(tmp_path, capsys):
    """18:23-25: a path that does not exist prints a line and returns None (before any device work)."""
    from b200restore import harness
    assert harness.evaluate_model(None, tmp_path / "nope", "Unified Restored") is None
    assert "Skipping Unified Restored: Path does not exist" in capsys.readouterr().out
    assert harness.benchmark_table(None, {"A": tmp_path / "nope"}, verbose=False) == {}
